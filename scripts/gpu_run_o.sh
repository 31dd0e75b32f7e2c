#!/bin/bash
# round-2 GPU batch O: start-vector fix, leaf sizes, tests
mkdir -p gpurun_out/o
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x -k "smooth_start or compute_modes_matches or golden" > gpurun_out/o/pytest.log 2>&1
tail -4 gpurun_out/o/pytest.log
timeout 300 python scripts/v0_probe.py 250 2>&1 | head -4 > gpurun_out/o/v0_probe.log
cat gpurun_out/o/v0_probe.log
timeout 400 python scripts/leaf_probe.py 500 8 12 16 24 32 > gpurun_out/o/leaf_probe.log 2>&1
cat gpurun_out/o/leaf_probe.log
