timeout 600 python -m pytest tests/test_gjk_gpu.py tests/test_epa_gpu.py tests/test_mpr_gpu.py tests/test_pipeline_gpu.py -x -q 2>&1 | grep -E "^E|passed|failed" | head
timeout 300 python scripts/r02_dev.py pipe 2>&1 | grep -E "epa|EPA|C5 shapes|gjk"
timeout 300 python bench.py --only mix --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120
