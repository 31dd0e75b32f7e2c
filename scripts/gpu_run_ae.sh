#!/bin/bash
# incidence-based pattern build of the FE space: parity suite (pattern bit-exact vs scipy / oracle), bench
mkdir -p gpurun_out/ae
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x > gpurun_out/ae/pytest.log 2>&1
tail -3 gpurun_out/ae/pytest.log
FW_TRACE=1 timeout 300 python scripts/assemble_only.py 500 1 2>&1 | grep -E "symbolic|rep 0" > gpurun_out/ae/symbolic_trace.log
cat gpurun_out/ae/symbolic_trace.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ae/bench.json 2> gpurun_out/ae/bench.err
tail -1 gpurun_out/ae/bench.err
FW_SYMBOLIC_FULL_SORT=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | tail -1
