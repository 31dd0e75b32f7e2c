#!/bin/bash
# round-2 GPU batch AE: which Schur-complement GEMM on the small-front levels (SIMT 64x64 with batched epilogue vs DMMA 64 / 128 tiles)
mkdir -p gpurun_out/ae
cd "$(dirname "$0")/.."
for v in "FW_X=0" "FW_BIG_GEMM_MIN=120" "FW_BIG_GEMM_MIN=200" "FW_BIG_GEMM_MIN=300" "FW_DMMA64_MAX=400" "FW_DMMA64_MAX=2000" "FW_BLOCKED_MIN_NS=65"; do
  echo "== $v" | tee -a gpurun_out/ae/variants.log
  env $v timeout 200 python scripts/factor_reps.py 500 2>&1 | tail -1 | tee -a gpurun_out/ae/variants.log
done
