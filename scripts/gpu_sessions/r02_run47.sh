timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^E|passed|failed" | head
timeout 300 python scripts/gjk_c1.py 2>&1 | tail -2
timeout 300 python scripts/r02_dev.py pipe 2>&1 | grep -E "gjk|C5 shapes"
timeout 300 python scripts/gjk_per_type.py 2>&1 | tail -12
