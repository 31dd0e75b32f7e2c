#!/bin/bash
# round-2 final batch: full parity suite, smoke, default bench (with CPU baseline), reference arm, sweep + geometry workloads
mkdir -p gpurun_out/final
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/final/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/final/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final/pytest.log
tail -4 gpurun_out/final/pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/final/smoke.log 2>&1
tail -1 gpurun_out/final/smoke.log
timeout 900 python bench.py > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err
echo "bench rc=$?" >> gpurun_out/final/bench.err
tail -2 gpurun_out/final/bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final/bench_reference.json 2> gpurun_out/final/bench_reference.err
echo "reference rc=$?" >> gpurun_out/final/bench_reference.err
tail -3 gpurun_out/final/bench_reference.err
timeout 300 python bench.py --workload sweep --steps 1 --warmup 1 --points-per-rank 4 > gpurun_out/final/bench_sweep.json 2> gpurun_out/final/bench_sweep.err
timeout 400 python bench.py --workload geometry --steps 1 --warmup 1 --points-per-rank 4 > gpurun_out/final/bench_geometry.json 2> gpurun_out/final/bench_geometry.err
python - <<PY
import json
for f in ("bench", "bench_reference", "bench_sweep", "bench_geometry"):
    try:
        d = json.loads(open("gpurun_out/final/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("unit"), d.get("ms_per_step"), d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
