#!/bin/bash
# mid-M: time every solution (token-tile width) per shape and M; which one should be the default?
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/midm; mkdir -p $OUT
for nk in "10240 8192" "8192 8192" "57344 8192" "8192 28672"; do set -- $nk
  for m in 17 24 32 48 64 96 128 192 256 384 512; do
    echo "== n=$1 k=$2 m=$m"
    timeout 120 tools/bench_matmul -m $m -n $1 -k $2 -atype bf16 -ctype bf16 -btype nvfp4 -warmup 3 -repeat 15 -algo tune 2>&1 | grep "^Matmul" | head -5 | sed -e "s/Backend: petit, batch: 1, //" -e "s/Matmul /  /"
  done
done > $OUT/tune.log 2>&1
tail -5 $OUT/tune.log; wc -l $OUT/tune.log
