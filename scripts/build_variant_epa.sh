#!/bin/bash
# scripts/build_variant_epa.sh NAME [-Dflag ...]: scripts/lib_epaNAME.so = library with epa.cu built under extra flags
set -e
name=$1; shift
cd "$(dirname "$0")/.."
python -m distance3d_b200.build > /dev/null
mkdir -p gpurun_out/d3dvar
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC "$@" \
  -c distance3d_b200/csrc/epa.cu -o gpurun_out/d3dvar/epa_$name.o
objs=$(ls distance3d_b200/build/*.o | grep -v '/epa.o')
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o scripts/lib_epa$name.so $objs gpurun_out/d3dvar/epa_$name.o
echo scripts/lib_epa$name.so
