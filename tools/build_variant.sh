#!/bin/bash
# Build a kernel variant of libhopedg.so for A/B timing on the GPU box:  tools/build_variant.sh <name> <extra nvcc flags...>
# -> hopefoam_b200/variants/libhopedg_<name>.so ; run with HDG_LIB_PATH=hopefoam_b200/variants/libhopedg_<name>.so python bench.py ...
set -e
name=$1; shift
cd "$(dirname "$0")/../hopefoam_b200/csrc"
mkdir -p ../variants build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c dg_kernels.cu -o build/dg_kernels_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libhopedg_$name.so build/ref_element.o build/mesh.o build/dg_kernels_$name.o build/dg_advect_tma.o build/dg_limiter.o build/hopedg.o -lcudart
echo built ../variants/libhopedg_$name.so
