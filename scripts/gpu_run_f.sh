#!/bin/bash
# round-2 GPU batch F: new tests (refinement loop, coupler sweep), graph-enabled bench, sweep + geometry workloads,
mkdir -p gpurun_out/f
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x > gpurun_out/f/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f/pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/f/bench.json 2> gpurun_out/f/bench.err
echo "bench rc=$?" >> gpurun_out/f/bench.err
timeout 300 python bench.py --workload sweep --steps 1 --warmup 1 --points-per-rank 4 > gpurun_out/f/bench_sweep.json 2> gpurun_out/f/bench_sweep.err
timeout 400 python bench.py --workload geometry --steps 1 --warmup 1 --points-per-rank 4 > gpurun_out/f/bench_geometry.json 2> gpurun_out/f/bench_geometry.err
timeout 400 python scripts/config3_probe.py 707 > gpurun_out/f/config3_probe.log 2>&1
timeout 400 python scripts/scale_probe.py 1000 > gpurun_out/f/scale_probe_1000.log 2>&1
timeout 200 python scripts/nrhs_probe.py 500 > gpurun_out/f/nrhs_probe.log 2>&1
