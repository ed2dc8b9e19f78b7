set -x
python -m pytest tests/test_broad_phase_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -4
for v in "" scripts/lib_bl4.so scripts/lib_bl16.so scripts/lib_bl32.so; do echo "== ${v:-default (8)}"; D3D_B200_LIB=$v python scripts/r02_dev.py bvh 2>&1 | grep -E "build|mode|two-pass"; done | tee gpurun_out/r02_bvh_block_sweep.txt
D3D_B200_LIB=scripts/lib_bl16.so python -m pytest tests/test_broad_phase_gpu.py -m gpu -x -q 2>&1 | tail -2
