#!/bin/bash
# Per-stage timeline of CTA 0 of the decode kernel (needs the -DPETIT_DEBUG_HOOKS variant:
#   tools/build_variant.sh hooks . -DPETIT_DEBUG_HOOKS
# The hooks cost ~20 % in the hot loops, so read the columns relative to each other):
#   w_issue act_issue | dq: full seen, a_empty seen, stores done | mma: a_full seen, committed
# Output: gpurun_out/stage_trace/<gemm>.log
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/stage_trace
mkdir -p $OUT
for s in ${1:-qkv gate_up}; do
  LD_LIBRARY_PATH=$PWD/variants/hooks PETIT_TRACE=1 PETIT_TRACE_STAGES=1 timeout 60 \
    tools/gemm_bench nv bf16 20 $s ${2:-16} > $OUT/$s.log 2>&1
  tail -30 $OUT/$s.log
done
