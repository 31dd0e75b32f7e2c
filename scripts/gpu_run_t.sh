#!/bin/bash
# round-2 GPU batch T: ncu --set full of one whole substitution pair and of the SpMV (CSV pages exported on the box), tests, bench
mkdir -p gpurun_out/t
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x -k "lu_solve or config2_at_scale or compute_modes_matches or sweep_reuses" > gpurun_out/t/pytest_lu.log 2>&1
tail -3 gpurun_out/t/pytest_lu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/t/bench.json 2> gpurun_out/t/bench.err
tail -1 gpurun_out/t/bench.err
timeout 900 ncu --set full --clock-control none -k regex:"k_fwd|k_bwd|k_tinv_fwd|k_tinv_matvec|k_tinv_bwd|k_upd_reduce|k_lu_permute" -c 90 -o /tmp/ncu_solve_pair python scripts/solve_once.py 500 > gpurun_out/t/ncu_pair.log 2>&1
ncu -i /tmp/ncu_solve_pair.ncu-rep --page raw --csv > gpurun_out/t/ncu_solve_pair_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:"k_spmv|k_op_apply" -c 3 -o /tmp/ncu_spmv python scripts/assemble_only.py 500 1 > gpurun_out/t/ncu_spmv.log 2>&1
ncu -i /tmp/ncu_spmv.ncu-rep --page raw --csv > gpurun_out/t/ncu_spmv_raw.csv 2>/dev/null
ls -la gpurun_out/t
