timeout 600 python -m pytest tests/test_epa_gpu.py tests/test_gjk_gpu.py tests/test_pipeline_gpu.py -x -q 2>&1 | grep -E "^E|passed|failed" | head
timeout 300 python scripts/epa_thread_dev.py c5 2>&1 | tail -1
timeout 300 python scripts/r02_dev.py pipe 2>&1 | grep -E "epa|C5 shapes"
