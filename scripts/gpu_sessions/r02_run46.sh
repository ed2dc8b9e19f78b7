echo default; timeout 300 python scripts/r02_dev.py pipe 2>&1 | grep -E "gjk"
echo mask3f; D3D_B200_LIB=scripts/lib_gm3f.so timeout 300 python scripts/r02_dev.py pipe 2>&1 | grep -E "gjk"
