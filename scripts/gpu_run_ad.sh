#!/bin/bash
# round-2 GPU batch AD: triangular fetch of the inverse diagonal blocks; forward fusion threshold
mkdir -p gpurun_out/ad
cd "$(dirname "$0")/.."
i=0
for v in "FW_X=0" "FW_FUSE_MIN_NS=0" "FW_FUSE_MIN_NS=48" "FW_TINV_MIN_NS=300" "FW_TINV_MIN_NS=400"; do
  i=$((i+1))
  echo "== $v" | tee -a gpurun_out/ad/variants.log
  env $v timeout 200 python scripts/nrhs_probe.py 500 > gpurun_out/ad/probe_$i.log 2>&1
  grep "fw-solve" gpurun_out/ad/probe_$i.log | awk '{print $3, $4}' | sort | awk '{a[$1]=a[$1]" "$2} END {for (k in a) print "nrhs", k, a[k]}' | tee -a gpurun_out/ad/variants.log
done
timeout 600 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x -k "lu_solve or schedules or config2" > gpurun_out/ad/pytest.log 2>&1
tail -3 gpurun_out/ad/pytest.log
