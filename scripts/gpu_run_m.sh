#!/bin/bash
# round-2 GPU batch M: 64x64 tensor-core GEMM tiles, gather/sequential interchanges by level, smooth start vector: tests, traces, bench
mkdir -p gpurun_out/m
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x > gpurun_out/m/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/m/pytest.log
tail -4 gpurun_out/m/pytest.log
for cfg in "" "FW_BIG_GEMM_MIN=384 FW_BIG_GEMM_MIN_OUTER=384" "FW_BIG_GEMM_MIN=100" "FW_BIG_GEMM_MIN_OUTER=200" "FW_DMMA64_MAX=600" "FW_DMMA64_MAX=1500" "FW_SWAP_GATHER_MAX_FRONTS=64" "FW_SWAP_GATHER_MAX_FRONTS=4096"; do
  tag=$(echo "$cfg" | tr ' =' '__'); [ -z "$tag" ] && tag=default
  env $cfg FW_FACTOR_TRACE=2 timeout 300 python scripts/trace_solve.py 500 2>&1 | grep -E "factor lev|residual" > gpurun_out/m/trace_$tag.log
  echo "$tag: panels $(grep panels gpurun_out/m/trace_$tag.log | awk '{s+=$(NF-1)} END {print s}') schur $(grep schur gpurun_out/m/trace_$tag.log | awk '{s+=$(NF-1)} END {print s}') $(grep residual gpurun_out/m/trace_$tag.log)"
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/m/bench.json 2> gpurun_out/m/bench.err
tail -1 gpurun_out/m/bench.err
timeout 300 python scripts/v0_probe.py 250 2>&1 | head -4 > gpurun_out/m/v0_probe.log
cat gpurun_out/m/v0_probe.log
