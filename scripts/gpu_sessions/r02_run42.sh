python -m pytest tests/test_epa_gpu.py -x -q 2>&1 | grep -E "^E|passed|failed" | head
echo prefetch; python scripts/epa_thread_dev.py c5 2>&1 | tail -1
echo no prefetch; D3D_B200_LIB=scripts/lib_epanopf.so python scripts/epa_thread_dev.py c5 2>&1 | tail -1
