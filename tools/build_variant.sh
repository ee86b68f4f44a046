#!/bin/bash
# Build a kernel variant of libhopedg.so for A/B timing on the GPU box:
#   tools/build_variant.sh <name> <source.cu> <extra nvcc flags...>     (source = dg_kernels.cu | dg_euler_split.cu | dg_limiter.cu | dg_advect_tma.cu)
# -> hopefoam_b200/variants/libhopedg_<name>.so ; run with HDG_LIB_PATH=hopefoam_b200/variants/libhopedg_<name>.so python bench.py ...
set -e
name=$1; src=$2; shift 2
cd "$(dirname "$0")/../hopefoam_b200/csrc"
mkdir -p ../variants build
base=${src%.cu}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c $src -o build/${base}_var_$name.o
objs=""
for o in ref_element mesh dg_kernels dg_euler_split dg_advect_tma dg_limiter hopedg; do
  if [ "$o" = "$base" ]; then objs="$objs build/${base}_var_$name.o"; else objs="$objs build/$o.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libhopedg_$name.so $objs -lcudart -ldl
echo built ../variants/libhopedg_$name.so
