#!/bin/bash
mkdir -p gpurun_out/r
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x > gpurun_out/r/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r/pytest.log
tail -4 gpurun_out/r/pytest.log
for cfg in "" "FW_TINV_MIN_NS=370"; do
  echo "== $cfg" >> gpurun_out/r/probe.log
  env $cfg FW_FACTOR_TRACE=1 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "residual|selective" >> gpurun_out/r/probe.log
  env $cfg timeout 200 python scripts/nrhs_probe.py 500 2>&1 | grep "fw-solve" | sed -n '3,5p;9,10p' >> gpurun_out/r/probe.log
done
cat gpurun_out/r/probe.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r/bench.json 2> gpurun_out/r/bench.err
tail -1 gpurun_out/r/bench.err
timeout 300 python bench.py --workload sweep --steps 1 --warmup 1 --points-per-rank 4 > gpurun_out/r/bench_sweep.json 2> gpurun_out/r/bench_sweep.err
python -c "
import json; d=json.load(open('gpurun_out/r/bench_sweep.json')); print('sweep', d['value'], d['last_point_stages_ms']['lu_factor_ms'], d['last_point_stages_ms']['arnoldi_ms'])"
