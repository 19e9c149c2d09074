#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck, initcheck) over the native selftest: every
# token-tile variant, ragged M, partial n-tiles, split (stream-K) tiles.  Output: gpurun_out/sanitize/
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/sanitize; mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --print-limit 20 tests/native/selftest > $OUT/$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SELFTEST|Internal|hazard" $OUT/$tool.log | head -5
done
