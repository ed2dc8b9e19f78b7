set -x
python -m pytest tests/test_norm_gpu.py tests/test_epa_gpu.py tests/test_gjk_gpu.py -x -q 2>&1 | tail -3
D3D_B200_LIB=scripts/lib_epaprof.so python scripts/epa_profile.py 2>&1 | tee gpurun_out/r02_epa_phases.txt
python scripts/r02_dev.py epa pipe 2>&1 | grep -E "epa|EPA|C5 shapes|gjk"
python scripts/gjk_c1.py 2>&1 | tail -2
