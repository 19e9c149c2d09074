#!/bin/bash
# One gpurun call: A/B of library variants (variants/<name>/libpetit_b200.so, see
# tools/build_variant.sh) on the four 70B decode GEMMs, parity of the default build, per-CTA
# traces and a bench line.  Output: gpurun_out/$EXP/ (default exp3).
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/${EXP:-exp3}
mkdir -p $OUT
B=tools/gemm_bench
run4() { for s in qkv o gate_up down; do echo -n "$1 "; timeout 60 $B ${3:-nv} ${4:-bf16} 60 $s ${2:-16}; done; }

# fail fast if a variant of the kernel hangs (register-pool accounting of setmaxnreg)
for m in 16 64 128 1024; do
  timeout 40 $B nv bf16 3 qkv $m > $OUT/smoke_$m.log 2>&1 || { echo "smoke M=$m failed rc=$?"; cat $OUT/smoke_$m.log; exit 1; }
done
timeout 40 $B mx bf16 3 qkv 16 >> $OUT/smoke_16.log 2>&1 || { echo "smoke mx failed"; exit 1; }
timeout 40 $B nv f16 3 qkv 16 >> $OUT/smoke_16.log 2>&1 || { echo "smoke f16 failed"; exit 1; }
timeout 400 python -m pytest tests -m gpu -x -q --timeout 90 > $OUT/pytest_default.log 2>&1
echo "pytest default rc=$?" | tee -a $OUT/pytest_default.log
{
  for rep in 1 2; do
    for v in $(ls variants); do LD_LIBRARY_PATH=$PWD/variants/$v run4 "variant=$v"; done
    run4 "variant=cur"
  done
  run4 "variant=cur M=1" 1
  run4 "variant=cur M=64" 64
  run4 "variant=cur mx" 16 mx
  run4 "variant=cur f16" 16 nv f16
  LD_LIBRARY_PATH=$PWD/variants/base run4 "variant=base M=64" 64
  LD_LIBRARY_PATH=$PWD/variants/base run4 "variant=base mx" 16 mx
  LD_LIBRARY_PATH=$PWD/variants/base run4 "variant=base f16" 16 nv f16
  for s in qkv gate_up; do for m in 1024 4096; do echo -n "variant=cur "; $B nv bf16 20 $s $m; echo -n "variant=base "; LD_LIBRARY_PATH=$PWD/variants/base $B nv bf16 20 $s $m; done; done
} > $OUT/variants.log 2>&1
for s in qkv o gate_up down; do
  PETIT_TRACE2=1 PETIT_TRACE_DUMP=$OUT/percta_cur.csv timeout 60 $B nv bf16 40 $s 16
done > $OUT/trace_cur.log 2>&1
timeout 300 python bench.py --steps 300 --warmup 5 --no-details > $OUT/bench_cur.json 2> $OUT/bench_cur.err
tail -3 $OUT/pytest_default.log; grep -v "^  " $OUT/variants.log
python -c "
import json
for f in ('bench_cur',):
    d=json.loads(open('$OUT/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], [(p['gemm'],p['us']) for p in d['roofline']['per_launch']])
"
