#!/bin/bash
# 64-token tile organisation: A/B of variants/* against the working-tree build (M = 64, 48, 128)
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/chains; mkdir -p $OUT
B=tools/gemm_bench
{
for rep in 1 2; do
 for m in 64 128; do for s in gate_up qkv down; do
  echo -n "variant=cur  "; PETIT_FORCE_NTOK=64 timeout 60 $B nv bf16 20 $s $m
  for v in $(ls variants); do
    echo -n "variant=$v "; PETIT_FORCE_NTOK=64 LD_LIBRARY_PATH=$PWD/variants/$v timeout 60 $B nv bf16 20 $s $m
  done
 done; done
done
for v in $(ls variants); do
  echo "parity variant=$v"; LD_LIBRARY_PATH=$PWD/variants/$v timeout 300 tests/native/selftest | grep 'FAIL\|SELFTEST' | head -5
done
} 2>&1 | tee $OUT/bench.log
