#!/bin/bash
# instruction-footprint experiment: timing only (the no-finish variants return wrong closest points)
run() { echo "== $1 [$2]"; D3D_B200_LIB=$PWD/scripts/lib_$1.so D3D_SKIP_MIX=$3 python scripts/gjk_per_type.py $2 2>&1 | grep -E "Mpairs"; }
run base sphere ""
run nofinish sphere ""
run sph sphere 1
run sph_nofinish sphere 1
