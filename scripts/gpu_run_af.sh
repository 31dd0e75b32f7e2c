#!/bin/bash
# round-2 GPU batch AF: DINV chains with up to 256 threads (levels 12 / 11 of config 2 in the fused forward pass)
mkdir -p gpurun_out/af
cd "$(dirname "$0")/.."
i=0
for v in "FW_DINV_MAX_THREADS=128" "FW_DINV_MAX_THREADS=256"; do
  i=$((i+1))
  echo "== $v" | tee -a gpurun_out/af/variants.log
  env $v timeout 200 python scripts/nrhs_probe.py 500 > gpurun_out/af/probe_$i.log 2>&1
  grep "fw-solve" gpurun_out/af/probe_$i.log | awk '{print $3, $4}' | sort | awk '{a[$1]=a[$1]" "$2} END {for (k in a) print "nrhs", k, a[k]}' | tee -a gpurun_out/af/variants.log
done
timeout 300 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x -k "lu_solve or schedules" > gpurun_out/af/pytest.log 2>&1
tail -2 gpurun_out/af/pytest.log
