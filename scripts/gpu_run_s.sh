#!/bin/bash
# round-2 GPU batch S: ncu --set full of one whole substitution pair and of the SpMV; inverse pivot blocks down to 110 rows
mkdir -p gpurun_out/s
cd "$(dirname "$0")/.."
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fwd|k_bwd|k_tinv_fwd|k_tinv_matvec|k_tinv_bwd|k_upd_reduce|k_lu_permute" -c 120 -o gpurun_out/s/ncu_solve_pair python scripts/solve_once.py 500 > gpurun_out/s/ncu_pair.log 2>&1
tail -2 gpurun_out/s/ncu_pair.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_spmv|k_op_apply" -c 3 -o gpurun_out/s/ncu_spmv python scripts/assemble_only.py 500 1 > gpurun_out/s/ncu_spmv.log 2>&1
tail -1 gpurun_out/s/ncu_spmv.log
for cfg in "FW_TINV_MIN_NS=110"; do
  echo "== $cfg" >> gpurun_out/s/probe.log
  env $cfg FW_FACTOR_TRACE=1 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "residual|selective" >> gpurun_out/s/probe.log
  env $cfg timeout 200 python scripts/nrhs_probe.py 500 2>&1 | grep "fw-solve" | sed -n '3,5p' >> gpurun_out/s/probe.log
done
cat gpurun_out/s/probe.log
