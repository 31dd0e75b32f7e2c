#!/bin/bash
# round-2 GPU batch Z: GEMM epilogues with batched loads (factor), fused small-front kernel re-test
mkdir -p gpurun_out/z
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x -k "lu or solve or mode or eig" > gpurun_out/z/pytest.log 2>&1
tail -3 gpurun_out/z/pytest.log
FW_FACTOR_TRACE=2 timeout 300 python scripts/trace_solve.py 500 > gpurun_out/z/trace_solve.log 2>&1
grep -E "factor lev|lu factor|residual" gpurun_out/z/trace_solve.log | head -50
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/z/bench.json 2> gpurun_out/z/bench.err
tail -1 gpurun_out/z/bench.json | cut -c1-300
tail -1 gpurun_out/z/bench.err
FW_FRONT_SMALL=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/z/bench_fs.json 2> gpurun_out/z/bench_fs.err
tail -1 gpurun_out/z/bench_fs.err
