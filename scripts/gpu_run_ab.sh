#!/bin/bash
mkdir -p gpurun_out/ab
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x > gpurun_out/ab/pytest.log 2>&1
tail -3 gpurun_out/ab/pytest.log
for cfg in "" "FW_SPMV_BC_TPR=4"; do
  env $cfg timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ab/bench_$cfg.json 2> gpurun_out/ab/bench_$cfg.err
  echo "== $cfg"; tail -1 gpurun_out/ab/bench_$cfg.err
  python -c "
import json; d=json.load(open('gpurun_out/ab/bench_$cfg.json')); s=d['solver']; print('arnoldi - solves', s['t_arnoldi_ms']-s['t_solve_ms_total'], 'e2e ms', d['e2e']['ms_per_step'])"
done
