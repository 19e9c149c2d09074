#!/bin/bash
# Round-closing GPU run: parity, bench line, ncu launch list + full capture, two-launch traces,
# M sweep.  Ordered by importance; everything lands in gpurun_out/final/.
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/final
mkdir -p $OUT
B=tools/gemm_bench
timeout 300 python -m pytest tests -m gpu -x -q --timeout 90 > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 200 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-details > $OUT/ncu_launches.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:fp4_gemm -c 8 -f -o $OUT/layerset \
  python bench.py --steps 1 --warmup 1 --no-details > $OUT/ncu_full.log 2>&1
ncu -i $OUT/layerset.ncu-rep --page raw --csv > $OUT/ncu_full_layerset_raw.csv 2>/dev/null
for s in qkv o gate_up down; do
  PETIT_TRACE2=1 PETIT_TRACE_DUMP=$OUT/percta_70b.csv timeout 60 $B nv bf16 40 $s 16
done > $OUT/trace_decode.log 2>&1
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
{
  for m in 1 4 8 16 32 64 128 256 512 1024 2048 4096; do
    for s in qkv o gate_up down; do timeout 60 $B nv bf16 20 $s $m; done
  done
  for s in qkv o gate_up down qkv_tp8 o_tp8 gate_up_tp8 down_tp8; do timeout 60 $B nv f16 20 $s 16; done
  for s in qkv o gate_up down qkv_tp8 o_tp8 gate_up_tp8 down_tp8; do timeout 60 $B mx bf16 20 $s 16; done
  for s in qkv_tp8 o_tp8 gate_up_tp8 down_tp8; do timeout 60 $B nv bf16 20 $s 16; done
} > $OUT/m_sweep.log 2>&1
PETIT_TRACE2=1 timeout 60 $B nv bf16 20 gate_up 1024 > $OUT/trace_prefill.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:fp4_gemm -c 2 -f -o $OUT/prefill \
  $B nv bf16 1 gate_up 1024 > $OUT/ncu_prefill.log 2>&1
ncu -i $OUT/prefill.ncu-rep --page raw --csv > $OUT/ncu_full_prefill_raw.csv 2>/dev/null
rm -f $OUT/prefill.ncu-rep
tail -2 $OUT/pytest_gpu.log; tail -c 1500 $OUT/bench_n1.json | head -c 1500; echo; cat $OUT/smoke.log | tail -3; ls -la $OUT
