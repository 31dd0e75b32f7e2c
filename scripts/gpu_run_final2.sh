#!/bin/bash
# round-2 final batch (second build): full parity suite, smoke, default bench (with CPU baseline), reference arm, sweep + geometry
# workloads, DRAM traffic of one substitution pair (ncu, the metrics of the committed summary only) and one --set full capture of
# the two leaf-level chain kernels
mkdir -p gpurun_out/final2
cd "$(dirname "$0")/.."
O=gpurun_out/final2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1
tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
echo "bench rc=$?" >> $O/bench.err
tail -2 $O/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
echo "reference rc=$?" >> $O/bench_reference.err
timeout 300 python bench.py --workload sweep --steps 1 --warmup 1 --points-per-rank 4 > $O/bench_sweep.json 2> $O/bench_sweep.err
timeout 400 python bench.py --workload geometry --steps 1 --warmup 1 --points-per-rank 4 > $O/bench_geometry.json 2> $O/bench_geometry.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
timeout 240 ncu --metrics $M --clock-control none -k regex:"k_fwd|k_bwd|k_tinv_fwd|k_tinv_matvec|k_tinv_bwd|k_upd_reduce|k_lu_permute" -c 100 --csv --page raw --log-file $O/ncu_solve_pair_raw.csv python scripts/solve_once.py 500 > $O/ncu_pair.log 2>&1
echo "ncu pair rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"k_fwd_chain_dinv|k_bwd_chain_dinv" -c 2 -o /tmp/ncu_leaf_chains python scripts/solve_once.py 500 > $O/ncu_leaf.log 2>&1
echo "ncu leaf rc=$?"
ncu -i /tmp/ncu_leaf_chains.ncu-rep --page raw --csv > $O/ncu_leaf_chains_raw.csv 2>/dev/null
ncu -i /tmp/ncu_leaf_chains.ncu-rep --page details --csv > $O/ncu_leaf_chains_details.csv 2>/dev/null
python - <<PY
import json
for f in ("bench", "bench_reference", "bench_sweep", "bench_geometry"):
    try:
        d = json.loads(open("$O/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("unit"), d.get("ms_per_step"), d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
