#!/bin/bash
# kernel tuning sweep: each variant is a build of the same library with different -D flags
for so in scripts/lib_*.so; do
  echo "== $so"
  D3D_B200_LIB=$PWD/$so python scripts/gjk_per_type.py 2>&1 | grep -E "sphere-sphere|capsule-capsule|ellipsoid-box|mix"
done
