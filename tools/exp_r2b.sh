#!/bin/bash
# quick correctness + timing of the working-tree build: native selftest, decode timings, GPU pytest
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/${EXP:-r2b}; mkdir -p $OUT
timeout 300 tests/native/selftest > $OUT/selftest.log 2>&1; echo "selftest rc=$?"; tail -3 $OUT/selftest.log
B=tools/gemm_bench
for rep in 1 2; do for s in qkv o gate_up down; do echo -n "cur "; timeout 60 $B nv bf16 60 $s 16; done; done 2>&1 | tee $OUT/decode.log
for s in qkv o gate_up down; do echo -n "cur M=1 "; timeout 60 $B nv bf16 60 $s 1; done 2>&1 | tee -a $OUT/decode.log
for s in gate_up down; do echo -n "mx "; timeout 60 $B mx bf16 60 $s 16; echo -n "f16 "; timeout 60 $B nv f16 60 $s 16; done 2>&1 | tee -a $OUT/decode.log
if [ -n "$PYTEST" ]; then timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest.log; fi
