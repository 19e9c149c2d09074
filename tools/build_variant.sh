#!/bin/bash
# Build an experiment copy of libpetit_b200.so:
#   tools/build_variant.sh <name> [<src root>] [--patch FILE]... [nvcc flags...]
# e.g. tools/build_variant.sh c64 . -DPETIT_ACC_COLS_64=128 -DPETIT_NUMACC_64=1
#      tools/build_variant.sh mine . --patch my_experiment.patch
# A --patch is applied (patch -p0 from the source root) to a scratch copy of csrc/, never to
# the tree.  The copy of the library lands in variants/<name>/ (git-ignored, travels with
# gpurun); run any tool against it with LD_LIBRARY_PATH=variants/<name> (the tools and the
# torch extension use RUNPATH), e.g. PARITY=<name> tools/exp_quick.sh.
set -e
NAME=$1; SRC=${2:-.}; shift; shift || true
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$(cd "$SRC" && pwd)
OUT=$ROOT/variants/$NAME; mkdir -p $OUT/obj
PATCHES=(); FLAGS=()
while [ $# -gt 0 ]; do
  if [ "$1" = "--patch" ]; then PATCHES+=("$(cd "$(dirname "$2")" && pwd)/$(basename "$2")"); shift 2; else FLAGS+=("$1"); shift; fi
done
CSRC=$SRC/petit-kernel_b200/csrc
if [ ${#PATCHES[@]} -gt 0 ]; then
  SCRATCH=$(mktemp -d); mkdir -p $SCRATCH/petit-kernel_b200; cp -r $CSRC $SCRATCH/petit-kernel_b200/csrc
  for p in "${PATCHES[@]}"; do (cd $SCRATCH && patch -s -p0 < "$p"); done
  CSRC=$SCRATCH/petit-kernel_b200/csrc
fi
for f in fp4_gemm repack capi allreduce; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a "${FLAGS[@]}" -O3 -std=c++17 -lineinfo \
    -Xcompiler -fPIC -I $SRC/include -I $CSRC -c $CSRC/$f.cu -o $OUT/obj/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libpetit_b200.so $OUT/obj/*.o -lcudart
rm -rf $OUT/obj; echo built $OUT/libpetit_b200.so
