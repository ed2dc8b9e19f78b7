echo default; timeout 300 python scripts/epa_thread_dev.py c5 2>&1 | tail -1
echo tm3f; D3D_B200_LIB=scripts/lib_epatm3f.so timeout 300 python scripts/epa_thread_dev.py c5 2>&1 | tail -1
