#!/bin/bash
# round-2 GPU batch L: gather-form row interchanges (tests, trace), Schur GEMM thresholds, ncu --set full of the Schur GEMMs
mkdir -p gpurun_out/l
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x -k "lu_solve or config2_at_scale or compute_modes_matches or rectangle or bend or metallic" > gpurun_out/l/pytest_lu.log 2>&1
tail -3 gpurun_out/l/pytest_lu.log
for thr in 384 256 160 96; do
  FW_BIG_GEMM_MIN=$thr FW_FACTOR_TRACE=2 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "factor lev|residual|lu factor" > gpurun_out/l/trace_schur_$thr.log
done
grep -h "schur" gpurun_out/l/trace_schur_384.log | head -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_lu_gemm_dmma" -c 6 -o gpurun_out/l/ncu_dmma python scripts/factor_once.py 500 > gpurun_out/l/ncu_dmma.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_lu_swap_trsm|k_lu_outer_trsm|k_lu_extend_add" -c 9 -o gpurun_out/l/ncu_trsm python scripts/factor_once.py 500 > gpurun_out/l/ncu_trsm.log 2>&1
tail -2 gpurun_out/l/ncu_dmma.log
