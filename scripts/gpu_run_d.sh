#!/bin/bash
# round-2 GPU batch D: solve graph + cluster panel threshold sweep + kernel-level launch list of one factorisation
mkdir -p gpurun_out/d
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -k "lu_solve or config2_at_scale or compute_modes_matches or sweep_reuses or eigen_b200" > gpurun_out/d/pytest.log 2>&1
echo "rc=$?" >> gpurun_out/d/pytest.log
for mf in 8 16 32 64; do
  FW_CLUSTER_PANEL_MAX_FRONTS=$mf FW_FACTOR_TRACE=2 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "factor lev .* panels|residual" > gpurun_out/d/trace_cluster_$mf.log
done
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/d/bench.json 2> gpurun_out/d/bench.err
FW_NO_SOLVE_GRAPH=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/d/bench_nograph.json 2> gpurun_out/d/bench_nograph.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/d/launches_factor.csv python scripts/trace_solve.py 500 > gpurun_out/d/ncu_trace.log 2>&1
tail -3 gpurun_out/d/pytest.log
