#!/bin/bash
mkdir -p gpurun_out/y
cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_lu_front_small" -c 1 -o /tmp/ncu_fs python scripts/factor_once.py 500 > gpurun_out/y/ncu.log 2>&1
ncu -i /tmp/ncu_fs.ncu-rep --page raw --csv > gpurun_out/y/fs_raw.csv 2>/dev/null
ncu -i /tmp/ncu_fs.ncu-rep --page source --csv --print-source cuda > gpurun_out/y/fs_source.csv 2>/dev/null
ls -la gpurun_out/y
