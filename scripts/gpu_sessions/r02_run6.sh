set -x
python -m pytest tests/test_gjk_gpu.py tests/test_epa_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -5
for v in "" scripts/lib_pq0.so scripts/lib_pq1b3.so scripts/lib_pq1b4r4.so scripts/lib_pq1b4r12.so; do D3D_B200_LIB=$v python scripts/gjk_c1.py 2>&1 | tail -2; done | tee gpurun_out/r02_gjk_variants.txt
