#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/grouped; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest.log
timeout 300 python tools/exp_grouped.py 2>&1 | tee $OUT/bench.log
timeout 600 python bench.py --no-competitors > $OUT/bench_line.json 2> $OUT/bench_err.log; tail -c 600 $OUT/bench_line.json
