python -m pytest tests/test_epa_gpu.py -x -q 2>&1 | tail -2
echo default; python scripts/epa_thread_dev.py c5 2>&1 | tail -1
for v in b2 b3 r24 r8 v64; do echo $v; D3D_B200_LIB=scripts/lib_epa$v.so python scripts/epa_thread_dev.py c5 2>&1 | tail -1; done
