#!/bin/bash
# knock-out: no token-tile traffic (PETIT_DEBUG_FLAGS=4, hooks build; results are wrong on purpose) at M = 32 / 64 / 128
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/ko_tok; mkdir -p $OUT
{
for m in 16 32 64 128; do for s in gate_up qkv; do
  echo -n "hooks base      "; LD_LIBRARY_PATH=$PWD/variants/hooks timeout 60 tools/gemm_bench nv bf16 20 $s $m
  echo -n "hooks no-tokens "; PETIT_DEBUG_FLAGS=4 LD_LIBRARY_PATH=$PWD/variants/hooks timeout 60 tools/gemm_bench nv bf16 20 $s $m
done; done
} 2>&1 | tee $OUT/bench.log
