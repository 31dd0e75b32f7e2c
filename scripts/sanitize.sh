#!/bin/bash
# compute-sanitizer passes over one small end-to-end solve (memcheck, racecheck on shared memory, initcheck, synccheck):
# order 2, 2*12^2 triangles, 2 modes + H field + overlaps + estimator (smoke path).  Output: gpurun_out/sanitize/*.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/sanitize
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
    python scripts/sanitize_target.py > gpurun_out/sanitize/$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize/summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitize/$tool.log | tee -a gpurun_out/sanitize/summary.txt
done
