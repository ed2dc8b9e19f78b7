echo default; python scripts/epa_thread_dev.py c3 2>&1 | tail -1
echo v256; D3D_B200_LIB=scripts/lib_epav256.so D3D_EPA_KERNEL=thread python scripts/epa_thread_dev.py c3 2>&1 | tail -1
