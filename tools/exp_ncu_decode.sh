#!/bin/bash
# ncu --set full capture (with SASS-level sampling) of one decode GEMM of the working-tree build
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/${EXP:-ncu}; mkdir -p $OUT
S=${1:-gate_up}; M=${2:-16}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fp4_gemm -s 3 -c 1 -f -o $OUT/$S \
   tools/gemm_bench ${FMT:-nv} ${ATYPE:-bf16} 3 $S $M > $OUT/ncu_$S.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/ncu_$S.log; ls -la $OUT
