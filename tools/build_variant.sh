#!/bin/bash
# Build an experiment copy of libpetit_b200.so: tools/build_variant.sh <name> [<src root>] [nvcc flags...]
# The copy lands in variants/<name>/ (git-ignored, travels with gpurun); run any tool against it
# with LD_LIBRARY_PATH=variants/<name> (the tools and the torch extension use RUNPATH).
set -e
NAME=$1; SRC=${2:-.}; shift; shift || true
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/variants/$NAME; mkdir -p $OUT/obj
for f in fp4_gemm repack capi allreduce; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a "$@" -O3 -std=c++17 -lineinfo \
    -Xcompiler -fPIC -I $SRC/include -I $SRC/petit-kernel_b200/csrc \
    -c $SRC/petit-kernel_b200/csrc/$f.cu -o $OUT/obj/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libpetit_b200.so $OUT/obj/*.o -lcudart
rm -rf $OUT/obj; echo built $OUT/libpetit_b200.so
