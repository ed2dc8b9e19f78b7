set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for b in libccd hydroelastic; do python bench.py --only $b 2>&1 | tail -1; done
