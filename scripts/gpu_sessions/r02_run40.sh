python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^E|passed|failed" | head
python scripts/gjk_c1.py 2>&1 | tail -2
