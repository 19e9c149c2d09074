#!/bin/bash
# One gpurun call: A/B of library variants (variants/<name>/libpetit_b200.so, see
# tools/build_variant.sh) and of the PETIT_RAMP knob on the four 70B decode GEMMs, parity of
# the default build and of the candidate settings.  Output: gpurun_out/exp2/.
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/exp2
mkdir -p $OUT
B=tools/gemm_bench
run4() { for s in qkv o gate_up down; do echo -n "$1 "; timeout 60 $B nv bf16 60 $s ${2:-16}; done; }

timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_default.log 2>&1
echo "pytest default rc=$?" | tee -a $OUT/pytest_default.log
{
  for v in base g2; do LD_LIBRARY_PATH=$PWD/variants/$v run4 "variant=$v"; done
  run4 "variant=cur"
  for v in base g2; do LD_LIBRARY_PATH=$PWD/variants/$v run4 "variant=$v"; done
  run4 "variant=cur"
  run4 "variant=cur M=1" 1
  LD_LIBRARY_PATH=$PWD/variants/g2 run4 "variant=g2 M=1" 1
} > $OUT/variants.log 2>&1
{
  for r in 8,2 12,3 16,4 16,6 24,6; do PETIT_RAMP=$r run4 "cur ramp=$r"; done
  for r in 12,3 16,4 24,6; do PETIT_RAMP=$r LD_LIBRARY_PATH=$PWD/variants/g2 run4 "g2 ramp=$r"; done
} > $OUT/ramp.log 2>&1
for s in qkv o gate_up down; do
  PETIT_TRACE2=1 PETIT_TRACE_DUMP=$OUT/percta_cur.csv timeout 60 $B nv bf16 40 $s 16
done > $OUT/trace_cur.log 2>&1
for s in qkv o gate_up down; do
  LD_LIBRARY_PATH=$PWD/variants/g2 PETIT_TRACE2=1 PETIT_TRACE_DUMP=$OUT/percta_g2.csv timeout 60 $B nv bf16 40 $s 16
done > $OUT/trace_g2.log 2>&1
timeout 300 python bench.py --steps 300 --warmup 5 --no-details > $OUT/bench_cur.json 2> $OUT/bench_cur.err
LD_LIBRARY_PATH=$PWD/variants/g2 PETIT_RAMP=16,4 timeout 300 python bench.py --steps 300 --warmup 5 --no-details > $OUT/bench_g2_ramp.json 2> $OUT/bench_g2_ramp.err
LD_LIBRARY_PATH=$PWD/variants/g2 PETIT_RAMP=16,4 timeout 300 python -m pytest tests -m gpu -x -q -k "sweep or shape or 70b or determin or tile or tp or gate_up or reference" > $OUT/pytest_g2_ramp.log 2>&1
echo "pytest g2+ramp rc=$?" | tee -a $OUT/pytest_g2_ramp.log
tail -3 $OUT/pytest_default.log; grep -v "^  " $OUT/variants.log; grep -v "^  " $OUT/ramp.log; tail -2 $OUT/pytest_g2_ramp.log
python -c "
import json
for f in ('bench_cur','bench_g2_ramp'):
    d=json.loads(open('$OUT/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
"
