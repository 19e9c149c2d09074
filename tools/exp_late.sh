#!/bin/bash
# arrival-skew compensation of the stream-K cuts: A/B by PETIT_TILT_LATE (same binary)
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/late; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3 | tee $OUT/pytest.log
{
for rep in 1 2; do
 for t in 0 2 4 6; do
  for s in qkv o down gate_up; do
   echo -n "late=$t "; PETIT_TILT_LATE=$t timeout 120 tools/gemm_bench nv bf16 40 $s 16
  done
 done
done
for t in 0 3 4 5 6; do
  echo -n "late=$t bench: "; PETIT_TILT_LATE=$t timeout 300 python bench.py --no-details --steps 400 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], round(d['ms_per_step']*1e3, 2), [(p['gemm'], p['us']) for p in d['roofline']['per_launch']], d['clocks']['reasons'])"
done
for t in 0 4; do echo -n "late=$t "; PETIT_TILT_LATE=$t timeout 120 tools/gemm_bench mx bf16 40 gate_up 16; echo -n "late=$t "; PETIT_TILT_LATE=$t timeout 120 tools/gemm_bench nv f16n 40 gate_up 16; echo -n "late=$t "; PETIT_TILT_LATE=$t timeout 120 tools/gemm_bench nv bf16 40 gate_up 64; done
} 2>&1 | tee $OUT/bench.log
