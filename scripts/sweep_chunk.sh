#!/bin/bash
# refill policy sweep: global cursor (c0) vs warp-private chunks of the sorted order
for so in scripts/lib_c*.so; do
  echo "== $so"
  D3D_B200_LIB=$PWD/$so python scripts/gjk_per_type.py sphere 2>&1 | grep -E "mix"
  D3D_N=4194304 D3D_B200_LIB=$PWD/$so python scripts/gjk_per_type.py sphere 2>&1 | grep -E "mix  "
done
D3D_B200_LIB=$PWD/scripts/lib_c256.so python -m pytest tests/test_gjk_gpu.py -x -q 2>&1 | tail -1
