#!/bin/bash
mkdir -p gpurun_out/q
cd "$(dirname "$0")/.."
for cfg in "" "FW_NO_TINV_SOLVE=1" "FW_TINV_MIN_NS=370" "FW_TINV_MIN_NS=450"; do
  echo "== $cfg" >> gpurun_out/q/probe.log
  env $cfg FW_FACTOR_TRACE=1 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "residual|selective" >> gpurun_out/q/probe.log
  env $cfg timeout 200 python scripts/nrhs_probe.py 500 2>&1 | grep "fw-solve" | sed -n '3,5p;9,10p' >> gpurun_out/q/probe.log
done
cat gpurun_out/q/probe.log
FW_FACTOR_TRACE=0 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "fwd chain  lev [0-8] |bwd chain  lev [0-8] |residual" > gpurun_out/q/trace_solve.log
cat gpurun_out/q/trace_solve.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/q/bench.json 2> gpurun_out/q/bench.err
tail -1 gpurun_out/q/bench.err
