#!/bin/bash
# kernel tuning sweep: each scripts/lib_*.so is a build of the library with different -D flags
# (scripts/build_variant.sh); prints sphere-sphere / C1 mix / 6-type mix throughput for each
for so in scripts/lib_*.so; do
  echo "== $so"
  D3D_B200_LIB=$PWD/$so python scripts/gjk_per_type.py sphere 2>&1 | grep -E "Mpairs"
done
