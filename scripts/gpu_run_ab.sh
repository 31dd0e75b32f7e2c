#!/bin/bash
# round-2 GPU batch AB: substitution schedule knobs re-tuned after the occupancy-hint change (scripts/nrhs_probe.py 500 per variant)
mkdir -p gpurun_out/ab
cd "$(dirname "$0")/.."
i=0
for v in "FW_X=0" "FW_NO_DINV_CHAIN=1" "FW_FUSE_TDIV=1" "FW_FUSE_TDIV=4" "FW_FUSE_MAX_M=0" "FW_TINV_MIN_NS=60" "FW_TINV_MIN_NS=200" "FW_NO_TINV_SOLVE=1" "FW_TREE=1" "FW_SOLVE_PF_KB=64" "FW_TINV_ALL_LEVELS=1"; do
  i=$((i+1))
  echo "== $v" | tee -a gpurun_out/ab/variants.log
  env $v timeout 200 python scripts/nrhs_probe.py 500 > gpurun_out/ab/probe_$i.log 2>&1
  grep "fw-solve" gpurun_out/ab/probe_$i.log | awk '{print $3, $4}' | sort | awk '{a[$1]=a[$1]" "$2} END {for (k in a) print "nrhs", k, a[k]}' | tee -a gpurun_out/ab/variants.log
done
timeout 300 python scripts/leaf_probe.py 500 12 16 24 32 > gpurun_out/ab/leaf_probe.log 2>&1
cat gpurun_out/ab/leaf_probe.log | tail -5
