"""Print the interesting keys of a bench.py JSON line: `python scripts/bench_summary.py FILE` (or stdin when piped)."""
import json, sys
src = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
d = json.loads(src.strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"])
print(d["roofline"])
print(d["stages_ms"])
print(d["solver"])
if "cpu_baseline" in d: print(d["cpu_baseline"])
