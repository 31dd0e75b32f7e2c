#!/bin/bash
# round-2 GPU batch V: load-batching probe (ptxas sinks the factor loads of the substitution kernels), LU tests, solve probe, bench
mkdir -p gpurun_out/v
cd "$(dirname "$0")/.."
(cd scripts/microbench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/rowdot_probe rowdot_probe.cu && timeout 120 /tmp/rowdot_probe) > gpurun_out/v/rowdot_probe.log 2>&1
cat gpurun_out/v/rowdot_probe.log
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x -k "lu or solve or mode or eig" > gpurun_out/v/pytest.log 2>&1
tail -4 gpurun_out/v/pytest.log
timeout 300 python scripts/nrhs_probe.py 500 > gpurun_out/v/nrhs_probe.log 2>&1
tail -12 gpurun_out/v/nrhs_probe.log
FW_NO_TINV_SOLVE=1 timeout 300 python scripts/nrhs_probe.py 500 > gpurun_out/v/nrhs_probe_notinv.log 2>&1
tail -3 gpurun_out/v/nrhs_probe_notinv.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/v/bench.json 2> gpurun_out/v/bench.err
tail -1 gpurun_out/v/bench.json | cut -c1-600
tail -2 gpurun_out/v/bench.err
