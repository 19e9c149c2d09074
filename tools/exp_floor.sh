#!/bin/bash
# One gpurun call: parity of the default build, per-CTA two-launch traces of the four 70B
# decode GEMMs and a per-launch floor decomposition on synthetic shapes (no split / forced
# 2-way split / slope).  Everything lands in gpurun_out/exp1/; summarise the CSVs with
# tools/analyze_percta.py.  (profiles/r01_floor_and_tilt.log also holds the sweep of a range-tilt
# knob that this script once ran; the knob was measured slower and removed.)
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/exp1
mkdir -p $OUT
B=tools/gemm_bench
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.log 2>&1

timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_default.log 2>&1
echo "pytest default rc=$?" | tee -a $OUT/pytest_default.log

for s in qkv o gate_up down; do
  PETIT_TRACE2=1 PETIT_TRACE_DUMP=$OUT/percta_70b.csv timeout 60 $B nv bf16 40 $s 16
done > $OUT/trace_70b.log 2>&1

# floor decomposition: 148 n-tiles x {1,2,8,32} k-tiles = exactly one whole tile per CTA (no
# split-tile reduction); 74 n-tiles = every tile split over two CTAs that both finish last;
# 1280x8192 = the TP-8 qkv shard (10 tiles split 15 ways)
for s in 18944x256 18944x512 18944x2048 18944x8192 9472x512 9472x4096 9472x8192 1280x8192; do
  PETIT_TRACE2=1 PETIT_TRACE_DUMP=$OUT/percta_floor.csv timeout 60 $B nv bf16 40 $s 16
done > $OUT/trace_floor.log 2>&1
for s in 18944x256 18944x2048 9472x512; do
  PETIT_PDL=0 timeout 60 $B nv bf16 40 $s 16
done > $OUT/floor_nopdl.log 2>&1

tail -3 $OUT/pytest_default.log; grep -v "^  " $OUT/trace_floor.log
