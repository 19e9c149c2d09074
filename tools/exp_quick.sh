#!/bin/bash
# Quick A/B of variants/* against the default build on the four 70B decode GEMMs (+ parity of
# one variant given as $PARITY).  Output: gpurun_out/$EXP/.
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/${EXP:-exp4}
mkdir -p $OUT
B=tools/gemm_bench
run4() { for s in qkv o gate_up down; do echo -n "$1 "; timeout 60 $B ${3:-nv} ${4:-bf16} 60 $s ${2:-16}; done; }
{
  for rep in 1 2; do
    run4 "variant=cur"
    for v in $(ls variants); do LD_LIBRARY_PATH=$PWD/variants/$v run4 "variant=$v"; done
  done
} > $OUT/variants.log 2>&1
if [ -n "$PARITY" ]; then
  LD_LIBRARY_PATH=$PWD/variants/$PARITY timeout 300 python -m pytest tests -m gpu -x -q --timeout 90 > $OUT/pytest_$PARITY.log 2>&1
  echo "pytest $PARITY rc=$?" | tee -a $OUT/pytest_$PARITY.log
  tail -3 $OUT/pytest_$PARITY.log
fi
grep -v "^  " $OUT/variants.log
