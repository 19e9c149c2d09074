#!/bin/bash
# ncu --set full of gate_up at M = 16 / 32 / 64: which pipe limits the 32- and 64-token tiles?  (raw pages only)
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/ncu_midm; mkdir -p $OUT
for M in 16 32 64; do
  timeout 300 ncu --set full --clock-control none -k regex:fp4_gemm -s 3 -c 1 -f -o /tmp/gate_up_m$M \
     tools/gemm_bench nv bf16 3 gate_up $M > $OUT/ncu_m$M.log 2>&1
  ncu -i /tmp/gate_up_m$M.ncu-rep --page raw --csv > $OUT/gate_up_m${M}_raw.csv 2>/dev/null
  ncu -i /tmp/gate_up_m$M.ncu-rep --page details --csv > $OUT/gate_up_m${M}_details.csv 2>/dev/null
  echo "M=$M rc=$?"
done
ls -la $OUT
