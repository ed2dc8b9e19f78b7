echo default; python scripts/epa_thread_dev.py c3 c5 2>&1 | tail -2
for v in b2r8 b3r8 b4r4 b4r16 b3r16 b4r8v16; do echo $v; D3D_B200_LIB=scripts/lib_epa$v.so python scripts/epa_thread_dev.py c5 2>&1 | tail -1; done
