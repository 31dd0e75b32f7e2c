#!/bin/bash
# round-2 GPU batch X: occupancy hints (__launch_bounds__ min blocks) on the substitution kernels -- without one ptxas
# minimises registers (32) and keeps ONE factor load in flight per thread
mkdir -p gpurun_out/x
cd "$(dirname "$0")/.."
for combo in "0 0" "1 0" "2 0" "2 1" "2 2" "1 1"; do
  set -- $combo
  echo "== FW_UPD_VAR=$1 FW_CHAIN_VAR=$2" | tee -a gpurun_out/x/variants.log
  FW_UPD_VAR=$1 FW_CHAIN_VAR=$2 timeout 300 python scripts/nrhs_probe.py 500 > gpurun_out/x/probe_$1_$2.log 2>&1
  grep "fw-solve" gpurun_out/x/probe_$1_$2.log | sort -k4 -n | awk '{print $3, $4}' | sort | awk '{a[$1]=a[$1]" "$2} END {for (k in a) print "nrhs", k, a[k]}' | tee -a gpurun_out/x/variants.log
done
FW_UPD_VAR=2 FW_CHAIN_VAR=1 timeout 300 python scripts/trace_solve.py 500 > gpurun_out/x/trace_solve_2_1.log 2>&1
grep -E "update|chain|residual" gpurun_out/x/trace_solve_2_1.log | head -70
