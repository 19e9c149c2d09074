#!/bin/bash
# Round-2 first measurement: A/B of the prepared variants, M=64 for cl64, per-stage trace.
cd "${GRAFT_REPO_ROOT:-.}"
export EXP=r2a
OUT=gpurun_out/$EXP; mkdir -p $OUT
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $OUT/clocks.csv &
SMI=$!
tools/exp_quick.sh > $OUT/quick.out 2>&1
B=tools/gemm_bench
for v in cur cl64; do for s in qkv o gate_up down; do
  echo -n "variant=$v M=64 "; if [ $v = cur ]; then timeout 60 $B nv bf16 60 $s 64; else LD_LIBRARY_PATH=$PWD/variants/$v timeout 60 $B nv bf16 60 $s 64; fi
done; done > $OUT/m64.log 2>&1
tools/exp_stage_trace.sh "qkv gate_up" > $OUT/stage_trace.out 2>&1
kill $SMI
cat $OUT/quick.out | tail -60; cat $OUT/m64.log
