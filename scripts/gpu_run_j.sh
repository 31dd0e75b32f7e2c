#!/bin/bash
# round-2 GPU batch J: L2 prefetch budget of the fused small-front substitution kernels, start-vector probe
mkdir -p gpurun_out/j
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x -k "lu_solve" > gpurun_out/j/pytest_lu.log 2>&1
tail -3 gpurun_out/j/pytest_lu.log
for pf in 0 8 16 32 48 80 160; do
  echo "== FW_SOLVE_PF_KB=$pf" >> gpurun_out/j/pf_probe.log
  FW_SOLVE_PF_KB=$pf timeout 200 python scripts/nrhs_probe.py 500 2>&1 | grep "fw-solve" | sed -n '3,5p;8,10p' >> gpurun_out/j/pf_probe.log
done
cat gpurun_out/j/pf_probe.log
timeout 600 python scripts/v0_probe.py 250 > gpurun_out/j/v0_probe.log 2>&1
cat gpurun_out/j/v0_probe.log
