#!/bin/bash
# Round-closing GPU run: parity, bench line, ncu launch list + full capture, two-launch traces,
# mode sweep.  Ordered by importance; everything lands in gpurun_out/final/.
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/final
mkdir -p $OUT
B=tools/gemm_bench
timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
timeout 600 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference_arm.json 2> $OUT/bench_reference_arm.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-details > $OUT/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fp4_gemm -s 12 -c 4 -f -o $OUT/layerset \
  python bench.py --steps 1 --warmup 3 --no-details > $OUT/ncu_full.log 2>&1
ncu -i $OUT/layerset.ncu-rep --page raw --csv > $OUT/ncu_full_layerset_raw.csv 2>/dev/null
rm -f $OUT/layerset.ncu-rep   # (gpurun_out/ may not exceed 64 MiB; the raw page is what profiles/ keeps)
for s in qkv o gate_up down; do
  PETIT_TRACE2=1 PETIT_TRACE_DUMP=$OUT/percta_70b.csv timeout 60 $B nv bf16 40 $s 16
done > $OUT/trace_decode.log 2>&1
python tools/analyze_percta.py $OUT/percta_70b.csv > $OUT/percta_summary.txt 2>&1
{
  for s in qkv o gate_up down; do for a in "nv bf16" "nv f16" "nv f16n" "mx bf16"; do timeout 60 $B $a 30 $s 16; done; done
  for s in qkv_tp8 o_tp8 gate_up_tp8 down_tp8; do timeout 60 $B nv bf16 30 $s 16; done
  for m in 32 64 128 256 512; do for s in qkv o gate_up down; do timeout 60 $B nv bf16 20 $s $m; done; done
} > $OUT/mode_sweep.log 2>&1
tail -2 $OUT/pytest_gpu.log; tail -2 $OUT/smoke.log; head -c 600 $OUT/bench_n1.json; echo; cat $OUT/bench_reference_arm.json | head -c 400; echo; ls -la $OUT
