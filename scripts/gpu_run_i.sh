#!/bin/bash
# round-2 GPU batch I: dataflow bands of the triangular solves (tests, A/B solve timings, bench), ncu --set full of the register panel kernel
mkdir -p gpurun_out/i
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x -k "lu_solve or config2_at_scale or compute_modes_matches or sweep_reuses" > gpurun_out/i/pytest_lu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/i/pytest_lu.log
tail -5 gpurun_out/i/pytest_lu.log
FW_FACTOR_TRACE=0 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "fwd|bwd|residual|solve wall" > gpurun_out/i/trace_solve_bands.log
for cfg in "" "FW_NO_TREE=1" "FW_TREE_MIN_FRONTS=128" "FW_TREE_MIN_FRONTS=32" "FW_TREE_SMALL_M=200" "FW_TREE_SMALL_M=400" "FW_TREE_THREADS_B=256" "FW_TREE_THREADS_A=64" "FW_TREE_THREADS_A=256"; do
  echo "== $cfg" >> gpurun_out/i/nrhs_probe.log
  env $cfg timeout 200 python scripts/nrhs_probe.py 500 2>&1 | grep "fw-solve" | tail -n 7 >> gpurun_out/i/nrhs_probe.log
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/i/bench.json 2> gpurun_out/i/bench.err
tail -2 gpurun_out/i/bench.err
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x > gpurun_out/i/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/i/pytest.log
tail -5 gpurun_out/i/pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_lu_panel_regs" --launch-skip 100 -c 2 -o gpurun_out/i/ncu_panel_regs python scripts/factor_once.py 500 > gpurun_out/i/ncu_panel.log 2>&1
cat gpurun_out/i/nrhs_probe.log
