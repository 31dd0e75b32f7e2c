#!/bin/bash
mkdir -p gpurun_out/w
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x -k "lu_solve or config2_at_scale or compute_modes_matches" > gpurun_out/w/pytest_lu.log 2>&1
tail -3 gpurun_out/w/pytest_lu.log
timeout 200 python scripts/nrhs_probe.py 500 2>&1 | grep "fw-solve" | sed -n '3,5p;9,10p'
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/w/bench.json 2> gpurun_out/w/bench.err
tail -1 gpurun_out/w/bench.err
