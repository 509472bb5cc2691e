#!/bin/bash
# last check of the in-tree library: large-state kernel tests, smoke, the default bench line with its entry-point parity check
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "C5 or big or C4 or smoke or seed" 2>&1 | tail -2
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline --no-configs > gpurun_out/last_bench.json 2> gpurun_out/last_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/last_bench.json")); print("%.4e"%d["value"], "%.4e"%d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"], d.get("parity_check"))
PY
