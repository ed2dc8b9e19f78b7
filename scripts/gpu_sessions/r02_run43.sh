timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^E|passed|failed" | head
timeout 300 python scripts/gjk_c1.py 2>&1 | tail -2
timeout 300 python scripts/epa_thread_dev.py c5 2>&1 | tail -1
timeout 300 python bench.py --only epa --no-cpu-baseline 2>&1 | tail -1 | cut -c1-140
