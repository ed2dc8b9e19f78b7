#!/bin/bash
# scripts/build_variant.sh NAME [-Dflag ...]: builds scripts/lib_NAME.so = the library with
# gjk.cu compiled under extra flags (kernel tuning sweeps, see sweep_variants.sh)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
python -m distance3d_b200.build > /dev/null
mkdir -p gpurun_out/d3dvar
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC "$@" \
  -c ${SRC:-distance3d_b200/csrc}/gjk.cu -o gpurun_out/d3dvar/gjk_$name.o
objs=$(ls distance3d_b200/build/*.o | grep -v '/gjk.o')
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o scripts/lib_$name.so $objs gpurun_out/d3dvar/gjk_$name.o
echo scripts/lib_$name.so
