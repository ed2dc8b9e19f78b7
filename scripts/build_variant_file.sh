#!/bin/bash
# scripts/build_variant_file.sh NAME FILE.cu [-Dflag ...]: scripts/lib_NAME.so = the library with
# csrc/FILE.cu compiled under extra flags (kernel tuning sweeps)
set -e
name=$1; file=$2; shift; shift
cd "$(dirname "$0")/.."
python -m distance3d_b200.build > /dev/null
mkdir -p gpurun_out/d3dvar
base=$(basename $file .cu)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC "$@" \
  -c distance3d_b200/csrc/$file -o gpurun_out/d3dvar/${base}_$name.o
objs=$(ls distance3d_b200/build/*.o | grep -v "/$base.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o scripts/lib_$name.so $objs gpurun_out/d3dvar/${base}_$name.o
echo scripts/lib_$name.so
