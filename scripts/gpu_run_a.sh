#!/bin/bash
# round-2 GPU batch A: parity suite, bench, FP64 microbench, per-level LU trace, sweep workload, config-3/5 probes
mkdir -p gpurun_out/a
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/a/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/a/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a/pytest.log
timeout 600 python bench.py --steps 3 --warmup 2 > gpurun_out/a/bench.json 2> gpurun_out/a/bench.err
echo "bench rc=$?" >> gpurun_out/a/bench.err
(cd scripts/microbench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && timeout 120 ./fp64_peak) > gpurun_out/a/fp64_peak.log 2>&1
FW_FACTOR_TRACE=2 timeout 300 python scripts/trace_solve.py 500 > gpurun_out/a/trace_solve.log 2>&1
timeout 300 python bench.py --workload sweep --steps 1 --warmup 1 --points-per-rank 3 > gpurun_out/a/bench_sweep.json 2> gpurun_out/a/bench_sweep.err
timeout 400 python scripts/config3_probe.py 707 > gpurun_out/a/config3_probe.log 2>&1
timeout 400 python scripts/scale_probe.py 1000 > gpurun_out/a/scale_probe_1000.log 2>&1
tail -3 gpurun_out/a/pytest.log
