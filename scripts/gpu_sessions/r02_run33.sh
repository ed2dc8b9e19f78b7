python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^E|passed|failed" | head -20
python bench.py --only pipeline --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1500
python bench.py --only epa --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1200
