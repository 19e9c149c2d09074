#!/bin/bash
# one-shot vs two-shot fused all-reduce at this world size (same box)
cd "${GRAFT_REPO_ROOT:-.}"
N=${1:-4}; OUT=gpurun_out/tp_oneshot; mkdir -p $OUT
for rep in 1 2; do
for ts in 1 0; do
  PETIT_AR_TWO_SHOT=$ts timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29562 bench.py --gpus $N --steps 300 --warmup 5 > $OUT/bench_tp${N}_ts$ts.json 2> $OUT/bench_tp${N}_ts$ts.err
  python - <<PY
import json
for line in open("$OUT/bench_tp${N}_ts$ts.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print("tp$N two_shot=$ts", d["value"], "GB/s", round(d["ms_per_step"]*1e3,2), "us/step", [(p["gemm"], p["us"]) for p in d["roofline"]["per_launch"]], d.get("tp_check",{}).get("ok"))
PY
done
done
