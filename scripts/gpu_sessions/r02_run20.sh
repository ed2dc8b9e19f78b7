set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/r02_dev.py epa pipe 2>&1 | grep -E "epa|EPA|C5 shapes|gjk" | tee gpurun_out/r02_epa_seedfix.txt
python scripts/gjk_c1.py 2>&1 | tail -2
for b in selfcollision; do python bench.py --only $b --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200; done
