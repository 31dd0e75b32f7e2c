#!/bin/bash
# round-2 GPU batch G: register panel kernel + two-level blocking of the LU factorisation (A/B traces), tests, bench
mkdir -p gpurun_out/g
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x > gpurun_out/g/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/g/pytest.log
FW_FACTOR_TRACE=2 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "factor lev|residual|lu factor" > gpurun_out/g/trace_default.log
FW_NO_REGS_PANEL=1 FW_FACTOR_TRACE=2 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "factor lev|residual|lu factor" > gpurun_out/g/trace_noregs.log
FW_BLOCKED_MIN_NS=100000 FW_FACTOR_TRACE=2 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "factor lev|residual|lu factor" > gpurun_out/g/trace_noblock.log
FW_BLOCKED_MIN_NS=129 FW_FACTOR_TRACE=2 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "factor lev|residual|lu factor" > gpurun_out/g/trace_block129.log
FW_BIG_GEMM_MIN_OUTER=200 FW_FACTOR_TRACE=2 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "factor lev|residual|lu factor" > gpurun_out/g/trace_outer200.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/g/bench.json 2> gpurun_out/g/bench.err
timeout 300 python scripts/scale_probe.py 1000 > gpurun_out/g/scale_probe_1000.log 2>&1
timeout 300 python scripts/config3_probe.py 707 > gpurun_out/g/config3_probe.log 2>&1
tail -12 gpurun_out/g/pytest.log; tail -2 gpurun_out/g/bench.err; grep residual gpurun_out/g/trace_*.log
