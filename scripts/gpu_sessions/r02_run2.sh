set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/r02_dev.py bvh epa pipe 2>&1 | tee gpurun_out/r02_dev_v2.txt
for b in 4 5 6; do echo "== EPA_BLOCKS_PER_SM=$b"; D3D_B200_LIB=scripts/lib_epaB$b.so python scripts/r02_dev.py epa pipe 2>&1 | grep -E "epa|EPA"; done | tee gpurun_out/r02_epa_occupancy.txt
