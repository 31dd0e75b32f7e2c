#!/bin/bash
# round-2 GPU batch E (re-entry): full parity suite, bench (both workloads), per-level LU traces, launch list, ncu --set full
mkdir -p gpurun_out/e
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/e/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider --durations=15 > gpurun_out/e/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/e/pytest.log
FW_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/e/bench.json 2> gpurun_out/e/bench.err
echo "bench rc=$?" >> gpurun_out/e/bench.err
FW_FACTOR_TRACE=2 timeout 300 python scripts/trace_solve.py 500 > gpurun_out/e/trace_solve.log 2>&1
timeout 300 python bench.py --workload sweep --steps 1 --warmup 1 --points-per-rank 3 > gpurun_out/e/bench_sweep.json 2> gpurun_out/e/bench_sweep.err
(cd scripts/microbench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && timeout 120 ./fp64_peak) > gpurun_out/e/fp64_peak.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/e/launches_n250.csv python bench.py --n 250 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/e/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_spmv|k_op_apply|k_fwd_chain_fused|k_bwd_chain_fused|k_lu_gemm_dmma|k_fwd_update|k_bwd_update" --launch-skip 40 -c 24 -o gpurun_out/e/ncu_solve_full python scripts/ncv_probe.py 500 4 20 > gpurun_out/e/ncu_full.log 2>&1
tail -25 gpurun_out/e/pytest.log
tail -3 gpurun_out/e/bench.err
