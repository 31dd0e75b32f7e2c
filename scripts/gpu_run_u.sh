#!/bin/bash
# multi-GPU batch: NCCL sweep test, weak-scaling bench (one cold config-2 solve per GPU per step), sweep workloads
N=${1:-2}
mkdir -p gpurun_out/u$N
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/u$N/smi.txt
timeout 300 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -k "two_gpus" > gpurun_out/u$N/pytest_nccl.log 2>&1
tail -2 gpurun_out/u$N/pytest_nccl.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/u$N/bench.json 2> gpurun_out/u$N/bench.err
tail -1 gpurun_out/u$N/bench.err
timeout 600 $TR --master-port 29512 bench.py --gpus $N --workload sweep --steps 1 --warmup 1 --points-per-rank 4 > gpurun_out/u$N/bench_sweep.json 2> gpurun_out/u$N/bench_sweep.err
timeout 600 $TR --master-port 29513 bench.py --gpus $N --workload geometry --steps 1 --warmup 1 --points-per-rank 4 > gpurun_out/u$N/bench_geometry.json 2> gpurun_out/u$N/bench_geometry.err
python - <<PY
import json
for f in ("bench", "bench_sweep", "bench_geometry"):
    try:
        d = json.loads(open("gpurun_out/u$N/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], d["value"], d["unit"], d["ms_per_step"], d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
