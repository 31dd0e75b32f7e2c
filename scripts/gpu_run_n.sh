#!/bin/bash
# round-2 GPU batch N: start-vector check, pinned result buffers (e2e), full tests
mkdir -p gpurun_out/n
cd "$(dirname "$0")/.."
FW_TRACE=1 timeout 300 python scripts/v0_probe.py 250 2>&1 | grep -E "smooth start|slab" | head -12 > gpurun_out/n/v0_probe.log
cat gpurun_out/n/v0_probe.log
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x > gpurun_out/n/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/n/pytest.log
tail -4 gpurun_out/n/pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n/bench.json 2> gpurun_out/n/bench.err
tail -1 gpurun_out/n/bench.err
python scripts/bench_summary.py gpurun_out/n/bench.json | head -3
