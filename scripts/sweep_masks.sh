#!/bin/bash
# how much does compiling out unused support cases buy? (instruction-fetch experiment)
run() { echo "== $1 [$2]"; D3D_B200_LIB=$PWD/scripts/lib_$1.so D3D_SKIP_MIX=$3 python scripts/gjk_per_type.py $2 2>&1 | grep -E "Mpairs"; }
run base sphere 1
run sphere sphere 1
run base box 1
run box box 1
run base capsule 1
run capsule capsule 1
run base sphere,capsule,box,ellipsoid,cylinder ""
run c1mask sphere ""
run refill24 sphere ""
run refill32 sphere ""
