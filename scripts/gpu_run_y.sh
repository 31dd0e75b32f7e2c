#!/bin/bash
# round-2 GPU batch Y: occupancy hints / batched loads across the substitution, Arnoldi, PCG and factor kernels -- full suite + bench
mkdir -p gpurun_out/y
cd "$(dirname "$0")/.."
timeout 300 python scripts/nrhs_probe.py 500 > gpurun_out/y/nrhs_probe.log 2>&1
grep "fw-solve" gpurun_out/y/nrhs_probe.log | tail -8
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/y/bench.json 2> gpurun_out/y/bench.err
tail -1 gpurun_out/y/bench.json | cut -c1-300
tail -2 gpurun_out/y/bench.err
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider --durations=8 > gpurun_out/y/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/y/pytest.log
tail -15 gpurun_out/y/pytest.log
timeout 300 python scripts/trace_solve.py 500 > gpurun_out/y/trace_solve.log 2>&1
grep -E "tinv|prep|matvec|residual" gpurun_out/y/trace_solve.log | head -40
