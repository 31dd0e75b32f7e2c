#!/bin/bash
# round-2 GPU batch AA: radix-sort load batching check + launch list of one config-2 solve (n = 500)
mkdir -p gpurun_out/aa
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x -k "coo or symbolic or pattern or analysis or lu_solve or config2" > gpurun_out/aa/pytest.log 2>&1
tail -3 gpurun_out/aa/pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/aa/bench.json 2> gpurun_out/aa/bench.err
tail -1 gpurun_out/aa/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/aa/launches_n500.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/aa/ncu_launches.log 2>&1
tail -2 gpurun_out/aa/ncu_launches.log | cut -c1-200
wc -l gpurun_out/aa/launches_n500.csv
