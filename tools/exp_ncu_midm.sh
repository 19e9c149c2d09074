#!/bin/bash
# ncu --set full of gate_up at mid / prefill M (which limiter: tensor pipe, L2->SM ingest, issue?)
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/ncu_midm; mkdir -p $OUT
for M in 128 256 1024; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:fp4_gemm -s 3 -c 1 -f -o $OUT/gate_up_m$M \
     tools/gemm_bench nv bf16 3 gate_up $M > $OUT/ncu_m$M.log 2>&1
  echo "M=$M ncu rc=$?"
done
PETIT_CLUSTER=0 timeout 300 ncu --set full --clock-control none -k regex:fp4_gemm -s 3 -c 1 -f -o $OUT/gate_up_m1024_nocl \
     tools/gemm_bench nv bf16 3 gate_up 1024 > $OUT/ncu_m1024_nocl.log 2>&1
for M in 128 256 1024; do tools/gemm_bench nv bf16 10 gate_up $M; PETIT_CLUSTER=0 tools/gemm_bench nv bf16 10 gate_up $M; done 2>&1 | tee $OUT/times.log
ls -la $OUT
