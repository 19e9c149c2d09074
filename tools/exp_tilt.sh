#!/bin/bash
# role-based stream-K cuts: parity, then A/B by PETIT_TILT_UNITS on the four decode shapes
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/tilt; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest.log
{
for rep in 1 2; do
 for t in 0 3 5 7; do
  for s in qkv o down gate_up; do
   echo -n "tilt=$t "; PETIT_TILT_UNITS=$t timeout 120 tools/gemm_bench nv bf16 40 $s 16
  done
 done
done
for t in 0 5; do echo -n "tilt=$t "; PETIT_TILT_UNITS=$t timeout 120 tools/gemm_bench nv bf16 40 qkv 1; echo -n "tilt=$t "; PETIT_TILT_UNITS=$t timeout 120 tools/gemm_bench nv bf16 40 down 64;  echo -n "tilt=$t "; PETIT_TILT_UNITS=$t timeout 120 tools/gemm_bench mx bf16 40 down 16; done
} 2>&1 | tee $OUT/bench.log
