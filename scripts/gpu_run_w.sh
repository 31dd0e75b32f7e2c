#!/bin/bash
# round-2 GPU batch W: DINV chains (pre-inverted diagonal blocks + cp.async double buffer) for the small-front levels
mkdir -p gpurun_out/w
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x -k "lu or solve or mode or eig" > gpurun_out/w/pytest.log 2>&1
tail -4 gpurun_out/w/pytest.log
timeout 300 python scripts/nrhs_probe.py 500 > gpurun_out/w/nrhs_probe.log 2>&1
grep "fw-solve" gpurun_out/w/nrhs_probe.log | tail -8
FW_NO_DINV_CHAIN=1 timeout 300 python scripts/nrhs_probe.py 500 > gpurun_out/w/nrhs_probe_nodinv.log 2>&1
grep "fw-solve" gpurun_out/w/nrhs_probe_nodinv.log | tail -8
timeout 300 python scripts/trace_solve.py 500 > gpurun_out/w/trace_solve.log 2>&1
grep -E "chain|residual|resid" gpurun_out/w/trace_solve.log | head -60
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/w/bench.json 2> gpurun_out/w/bench.err
tail -1 gpurun_out/w/bench.json | cut -c1-300
tail -2 gpurun_out/w/bench.err
