python -m pytest tests/test_epa_gpu.py tests/test_gjk_gpu.py tests/test_mpr_gpu.py -x -q 2>&1 | grep -E "^E|passed|failed" | head -20
python scripts/r02_dev.py epa 2>&1 | grep -E "epa|EPA|gjk"
python bench.py --only epa --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
python scripts/gjk_c1.py 2>&1 | tail -2
