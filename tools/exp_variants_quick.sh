#!/bin/bash
# A/B of variants/* against the working-tree build: gate_up and qkv at M=16 (timing only)
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/${EXP:-r2c}; mkdir -p $OUT
B=tools/gemm_bench
SHAPES=${SHAPES:-"gate_up qkv"}
{
for rep in 1 2; do
  for s in $SHAPES; do echo -n "variant=cur "; timeout 60 $B ${FMT:-nv} ${ATYPE:-bf16} 40 $s ${M:-16}; done
  for v in $(ls variants); do for s in $SHAPES; do echo -n "variant=$v "; LD_LIBRARY_PATH=$PWD/variants/$v timeout 60 $B ${FMT:-nv} ${ATYPE:-bf16} 40 $s ${M:-16}; done; done
done
} 2>&1 | tee $OUT/variants.log
