echo default; timeout 300 python bench.py --only epa --no-cpu-baseline 2>&1 | tail -1 | cut -c1-140
echo prefetchB; D3D_B200_LIB=scripts/lib_epapfb.so timeout 300 python bench.py --only epa --no-cpu-baseline 2>&1 | tail -1 | cut -c1-140
