#!/bin/bash
# round-2 GPU batch H: per-kernel list of one factorisation (ncu, factor kernels only)
mkdir -p gpurun_out/h
cd "$(dirname "$0")/.."
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_lu_" -c 1500 --csv --log-file gpurun_out/h/launches_factor.csv python scripts/factor_once.py 500 > gpurun_out/h/ncu.log 2>&1
tail -3 gpurun_out/h/ncu.log
