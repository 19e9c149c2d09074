#!/bin/bash
# per-CTA two-launch traces of the four 70B decode GEMMs with the working-tree build
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/${EXP:-trace4}; mkdir -p $OUT
for s in ${SHAPES:-qkv o gate_up down}; do
  PETIT_TRACE2=1 PETIT_TRACE_DUMP=$OUT/percta.csv timeout 60 tools/gemm_bench ${FMT:-nv} ${ATYPE:-bf16} 40 $s ${M:-16}
done > $OUT/trace.log 2>&1
python tools/analyze_percta.py $OUT/percta.csv > $OUT/percta_summary.txt 2>&1
cat $OUT/trace.log | head -120
