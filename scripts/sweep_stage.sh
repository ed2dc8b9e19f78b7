#!/bin/bash
# staging-buffer size of the single-pass LBVH query (shared memory per warp vs occupancy)
for so in scripts/lib_s*.so; do echo "== $so"; D3D_B200_LIB=$PWD/$so python scripts/bvh_bench.py 2>&1 | tail -2; done
