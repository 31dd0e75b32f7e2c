#!/bin/bash
# round-2 GPU batch C: DMMA Schur GEMM (LU tests, per-level factor trace at several thresholds), failing test of batch B
mkdir -p gpurun_out/c
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -k "lu_solve or mass_solve_direct or config2_at_scale or compute_modes_matches" > gpurun_out/c/pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c/pytest.log
for thr in 384 256 160; do
  FW_BIG_GEMM_MIN=$thr FW_FACTOR_TRACE=2 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "factor lev|residual|lu factor" > gpurun_out/c/trace_dmma_$thr.log
done
FW_NO_DMMA=1 FW_FACTOR_TRACE=2 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "factor lev|residual|lu factor" > gpurun_out/c/trace_nodmma.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/c/bench.json 2> gpurun_out/c/bench.err
tail -5 gpurun_out/c/pytest.log
