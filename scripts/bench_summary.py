import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"])
print(d["roofline"])
print(d["stages_ms"])
print(d["solver"])
if "cpu_baseline" in d: print(d["cpu_baseline"])
