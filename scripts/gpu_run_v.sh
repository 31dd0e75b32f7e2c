#!/bin/bash
# round-2 GPU batch V: L2-sized chunks of small fronts in the factorisation
mkdir -p gpurun_out/v
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x -k "lu_solve or config2_at_scale or compute_modes_matches" > gpurun_out/v/pytest_lu.log 2>&1
tail -3 gpurun_out/v/pytest_lu.log
for mb in 0 16 32 48 64 96; do
  FW_FACTOR_L2_MB=$mb FW_FACTOR_TRACE=1 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "lu factor: levels|residual" | tr '\n' ' ' | sed "s/^/L2_MB=$mb /" >> gpurun_out/v/probe.log
  echo >> gpurun_out/v/probe.log
done
cat gpurun_out/v/probe.log
FW_FACTOR_TRACE=2 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "factor lev (15|14|13|12|11|10) " > gpurun_out/v/trace_levels.log
cat gpurun_out/v/trace_levels.log
