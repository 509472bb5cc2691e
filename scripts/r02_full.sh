#!/bin/bash
# full GPU test suite + the default bench line
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/full_tests.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cat gpurun_out/full_tests.log; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_default.json"))
print(json.dumps({k:d[k] for k in d if k not in ("configs",)},indent=1)[:3500])
for k,v in d.get("configs",{}).items(): print(k, json.dumps(v)[:600])
PY
tail -5 gpurun_out/bench_default.err
