#!/bin/bash
# TP bench line(s) with the working-tree build (run with gpurun --gpus N)
cd "${GRAFT_REPO_ROOT:-.}"
N=${1:-2}; OUT=gpurun_out/tp_bench; mkdir -p $OUT
for mode in ${MODES:-fused}; do
  PETIT_TP_ALLREDUCE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29562 bench.py --gpus $N --steps ${STEPS:-300} --warmup 5 ${BENCH_ARGS} > $OUT/bench_tp${N}_$mode.json 2> $OUT/bench_tp${N}_$mode.err
  echo "bench tp$N $mode rc=$?"
  python - <<PY
import json
for line in open("$OUT/bench_tp${N}_$mode.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print("tp$N $mode", d["value"], "GB/s", round(d["ms_per_step"]*1e3,2), "us/step", [(p["gemm"], p["us"]) for p in d["roofline"]["per_launch"]], d.get("tp_check",{}).get("ok"))
PY
done
