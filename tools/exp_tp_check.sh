#!/bin/bash
# Multi-GPU validation of the peer all-reduce (run with gpurun --gpus N): stress test with
# data that changes every call in all four flag modes, the 2-GPU pytest, a TP bench line with
# tp_check, and (1 GPU) compute-sanitizer over the selftest.  Output: gpurun_out/tp_check/.
cd "${GRAFT_REPO_ROOT:-.}"
N=${1:-2}; ITERS=${2:-10000}
OUT=gpurun_out/tp_check; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus_$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
  --master-port 29551 tests/tp_peer_allreduce_check.py $ITERS $OUT/stress_world$N.json \
  > $OUT/stress_world$N.log 2>&1
echo "stress world=$N rc=$?" | tee -a $OUT/stress_world$N.log
tail -3 $OUT/stress_world$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
  --master-port 29552 bench.py --gpus $N --steps 200 --warmup 5 > $OUT/bench_tp$N.json 2> $OUT/bench_tp$N.err
echo "bench tp$N rc=$?"; tail -c 1500 $OUT/bench_tp$N.json
if [ "$N" = 2 ]; then
  timeout 600 python -m pytest tests -m gpu -x -q -k "peer_allreduce" > $OUT/pytest_peer.log 2>&1
  echo "pytest peer rc=$?"; tail -3 $OUT/pytest_peer.log
fi
