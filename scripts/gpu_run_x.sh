#!/bin/bash
# round-2 GPU batch X: fused small-front factorisation kernel
mkdir -p gpurun_out/x
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x -k "lu_solve or config2_at_scale or compute_modes_matches or rectangle or metallic or bend" > gpurun_out/x/pytest_lu.log 2>&1
tail -3 gpurun_out/x/pytest_lu.log
for cfg in "" "FW_NO_FRONT_SMALL=1" "FW_FRONT_SMALL_MAX_M=160" "FW_FRONT_SMALL_MAX_M=300" "FW_FRONT_SMALL_MAX_M=600"; do
  tag=$(echo "$cfg" | tr ' =' '__'); [ -z "$tag" ] && tag=default
  env $cfg FW_FACTOR_TRACE=2 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "factor lev|lu factor: levels|residual" > gpurun_out/x/trace_$tag.log
  echo "$tag: total $(grep 'factor lev' gpurun_out/x/trace_$tag.log | awk '{s+=$(NF-1)} END {print s}') $(grep residual gpurun_out/x/trace_$tag.log)"
done
grep -E "lev (15|14|13|12|11|10) " gpurun_out/x/trace_default.log
