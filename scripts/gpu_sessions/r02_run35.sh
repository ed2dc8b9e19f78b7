python -m pytest tests/test_epa_gpu.py tests/test_gjk_gpu.py -x -q 2>&1 | grep -E "^E|passed|failed" | head -20
python scripts/epa_thread_dev.py c3 c5 2>&1 | tail -2
python bench.py --only epa --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
