#!/bin/bash
# round-2 GPU batch B: device analysis vs host analysis, failing tests of batch A, bench
mkdir -p gpurun_out/b
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x -k "device_lu_analysis" > gpurun_out/b/pytest_analysis.log 2>&1
echo "rc=$?" >> gpurun_out/b/pytest_analysis.log
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/b/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b/pytest.log
FW_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/b/bench.json 2> gpurun_out/b/bench.err
echo "bench rc=$?" >> gpurun_out/b/bench.err
tail -3 gpurun_out/b/pytest_analysis.log; tail -15 gpurun_out/b/pytest.log
