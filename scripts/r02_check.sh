#!/bin/bash
# quick end-of-session check: GPU test suite, smoke, the default bench line and the reference arm as the driver runs them
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_driverlike.json 2> gpurun_out/r02_bench_driverlike.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_driverlike.json"))
print("value %.4e e2e %.4e frac %.3f launches %d"%(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"]), d["parity_check"], d["cpu_baseline"]["value"], d["roofline"]["traffic"])
for k,v in d["configs"].items(): print(k, "%.3e"%v.get("value",0), (v.get("cpu_baseline") or {}).get("value"), v.get("error"))
PY
