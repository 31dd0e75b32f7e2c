#!/bin/bash
# round-2 GPU batch K: shifted-column register panel, window panel variants: tests + factor trace + per-kernel list
mkdir -p gpurun_out/k
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x -k "lu_solve or config2_at_scale or compute_modes_matches or rectangle or bend or metallic" > gpurun_out/k/pytest_lu.log 2>&1
tail -3 gpurun_out/k/pytest_lu.log
FW_FACTOR_TRACE=2 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "factor lev|residual|lu factor" > gpurun_out/k/trace_default.log
cat gpurun_out/k/trace_default.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_lu_" -c 1500 --csv --log-file gpurun_out/k/launches_factor.csv python scripts/factor_once.py 500 > gpurun_out/k/ncu.log 2>&1
tail -2 gpurun_out/k/ncu.log
