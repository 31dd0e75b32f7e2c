#!/bin/bash
# round-2 GPU batch AG: asynchronous per-solve timing in the Arnoldi loop (no host sync after every substitution pair)
mkdir -p gpurun_out/ag
cd "$(dirname "$0")/.."
timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ag/bench.json 2> gpurun_out/ag/bench.err
tail -1 gpurun_out/ag/bench.err
timeout 200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x -k "eig or compute_modes or sweep_reuses or fall_back" > gpurun_out/ag/pytest.log 2>&1
tail -2 gpurun_out/ag/pytest.log
python - <<PY
import json
d = json.loads(open("gpurun_out/ag/bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["solver"]["t_solve_ms_total"], d["solver"]["n_op"])
PY
