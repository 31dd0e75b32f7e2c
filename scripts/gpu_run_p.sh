#!/bin/bash
# round-2 GPU batch P: selective inversion of the upper-level pivot blocks: tests, accuracy, solve time, bench
mkdir -p gpurun_out/p
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x -k "lu_solve or config2_at_scale or compute_modes_matches or rectangle or metallic or sweep_reuses" > gpurun_out/p/pytest_lu.log 2>&1
tail -4 gpurun_out/p/pytest_lu.log
for cfg in "" "FW_NO_TINV_SOLVE=1" "FW_TINV_MIN_NS=200" "FW_TINV_MIN_NS=60" "FW_TINV_MIN_NS=370"; do
  echo "== $cfg" >> gpurun_out/p/probe.log
  env $cfg FW_FACTOR_TRACE=1 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "residual|selective" >> gpurun_out/p/probe.log
  env $cfg timeout 200 python scripts/nrhs_probe.py 500 2>&1 | grep "fw-solve" | sed -n '3,5p;9,10p' >> gpurun_out/p/probe.log
done
cat gpurun_out/p/probe.log
FW_FACTOR_TRACE=0 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "fwd|bwd|residual" > gpurun_out/p/trace_solve.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/p/bench.json 2> gpurun_out/p/bench.err
tail -1 gpurun_out/p/bench.err
python scripts/bench_summary.py gpurun_out/p/bench.json | tail -2
