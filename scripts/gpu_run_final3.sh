#!/bin/bash
# round-2 final batch, last build: DRAM traffic of one substitution pair (ncu, summary written to profiles/ before the bench reads
# it), full parity suite, smoke, default bench (with CPU baseline), sweep + geometry workloads
mkdir -p gpurun_out/final3
cd "$(dirname "$0")/.."
O=gpurun_out/final3
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/smi.txt 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
timeout 240 ncu --metrics $M --clock-control none -k regex:"k_fwd|k_bwd|k_tinv_fwd|k_tinv_matvec|k_tinv_bwd|k_upd_reduce|k_lu_permute" -c 100 --csv --page raw --log-file $O/ncu_solve_pair_raw.csv python scripts/solve_once.py 500 > $O/ncu_pair.log 2>&1
echo "ncu pair rc=$?"
python scripts/ncu_raw_to_summary.py $O/ncu_solve_pair_raw.csv profiles/ncu_r02_final_solve_pair.csv 'ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_{long_scoreboard,barrier}_per_issue_active.ratio --clock-control none -k regex:"k_fwd|k_bwd|k_tinv_fwd|k_tinv_matvec|k_tinv_bwd|k_upd_reduce|k_lu_permute" -c 100 python scripts/solve_once.py 500 (FINAL round-2 build, scripts/gpu_run_final3.sh): every kernel of ONE forward/backward substitution pair at config 2 (no CUDA graph; cold-cache, serialised); sum of dram read+write = bench.py roofline.traffic'
cp profiles/ncu_r02_final_solve_pair.csv $O/ncu_r02_final_solve_pair.csv
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1
tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
echo "bench rc=$?" >> $O/bench.err
tail -2 $O/bench.err
timeout 300 python bench.py --workload sweep --steps 1 --warmup 1 --points-per-rank 4 > $O/bench_sweep.json 2> $O/bench_sweep.err
timeout 400 python bench.py --workload geometry --steps 1 --warmup 1 --points-per-rank 4 > $O/bench_geometry.json 2> $O/bench_geometry.err
FW_DMMA64_MAX=4000 timeout 200 python scripts/factor_reps.py 500 2>&1 | tail -1 > $O/factor_dmma64_4000.log
cat $O/factor_dmma64_4000.log
python - <<PY
import json
for f in ("bench", "bench_sweep", "bench_geometry"):
    try:
        d = json.loads(open("$O/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("unit"), d.get("ms_per_step"), d.get("e2e", {}).get("value"), d.get("roofline", {}).get("frac"), d.get("roofline", {}).get("traffic"))
    except Exception as e:
        print(f, "ERR", e)
PY
