#!/bin/bash
# dev: build monsoon_b200/libsb_b200_<name>.so with extra -D flags for the warp engine only (the thread-engine object is cached)
#   tools/build_variant.sh name [-DFLAG ...]      then   SB_LIB=monsoon_b200/libsb_b200_name.so python tools/...
set -e
cd "$(dirname "$0")/../monsoon_b200/csrc"
name=$1; shift
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off"
mkdir -p ../../build
if [ ! -f ../../build/sb_kernels.o ] || [ sb_kernels.cu -nt ../../build/sb_kernels.o ]; then nvcc $F -c sb_kernels.cu -o ../../build/sb_kernels.o; fi
nvcc $F "$@" -c sbw_kernels.cu -o ../../build/sbw_kernels_$name.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../libsb_b200_$name.so ../../build/sb_kernels.o ../../build/sbw_kernels_$name.o
echo built libsb_b200_$name.so
