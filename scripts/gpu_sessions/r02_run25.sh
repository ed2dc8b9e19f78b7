python -m pytest tests/test_epa_gpu.py -x -q 2>&1 | tail -4
python scripts/epa_thread_dev.py 2>&1 | tail -12
