echo default; timeout 300 python scripts/epa_thread_dev.py c5 2>&1 | tail -1
for v in v8q4 v8q8 v6q6; do echo $v; D3D_B200_LIB=scripts/lib_epa$v.so timeout 300 python scripts/epa_thread_dev.py c5 2>&1 | tail -1; done
