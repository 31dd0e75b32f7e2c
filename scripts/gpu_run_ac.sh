#!/bin/bash
# launch list of the final build (whole solve, 125k-triangle mesh: ncu costs ~0.15 s per launch)
mkdir -p gpurun_out/ac
cd "$(dirname "$0")/.."
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file /tmp/launches_n250.csv python bench.py --n 250 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ac/ncu_launches.log 2>&1
python - <<'PY'
import csv, re, collections
rows = [r for r in csv.reader(open('/tmp/launches_n250.csv')) if r and not r[0].startswith('==')]
hdr = rows[0]
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= iv: continue
    name = re.sub(r'\(.*', '', r[ik])
    v = float(r[iv].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[iu], 1.0)
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
with open('gpurun_out/ac/launches_summary.csv', 'w') as f:
    f.write('# ncu --metrics gpu__time_duration.sum --clock-control none -c 9000: python bench.py --n 250 --steps 1 --warmup 0 --no-cpu-baseline (125k triangles, FINAL round-2 build; graph kernel nodes listed individually); total %.1f us in %d launches\n' % (tot, sum(v[0] for v in agg.values())))
    f.write('kernel,launches,total_us,share\n')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('"%s",%d,%.1f,%.4f\n' % (k, v[0], v[1], v[1] / tot))
print(open('gpurun_out/ac/launches_summary.csv').read()[:3000])
PY
