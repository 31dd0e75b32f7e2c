#!/bin/bash
# round-2 GPU batch AC: occupancy hints of the row-interchange/TRSM kernel and the SIMT GEMM of the numeric factorisation
mkdir -p gpurun_out/ac
cd "$(dirname "$0")/.."
for v in "FW_X=0" "FW_SWAP_MINB=4" "FW_SWAP_MINB=6" "FW_GEMM_MINB=4" "FW_SWAP_GATHER_MAX_FRONTS=1000000" "FW_SWAP_GATHER_MAX_FRONTS=0" "FW_SWAP_GATHER_MAX_FRONTS=4096"; do
  echo "== $v" | tee -a gpurun_out/ac/variants.log
  env $v timeout 200 python scripts/factor_reps.py 500 2>&1 | tail -1 | tee -a gpurun_out/ac/variants.log
done
