#!/bin/bash
# full GPU session: all parity tests + headline bench lines
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
for st in stored recompute; do for ls in gmres neumann; do
  timeout 300 python bench.py --linsolve $ls --stage $st --no-cpu-baseline --steps 3 > gpurun_out/bench_${st}_${ls}.json 2> gpurun_out/bench_${st}_${ls}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${st}_${ls}.json"))
    print("$st $ls", "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], d["kernel_ms"], "fp64 %.3f" % d["fp64"]["frac"])
except Exception as ex:
    print("$st $ls failed", ex); print(open("gpurun_out/bench_${st}_${ls}.err").read()[-800:])
PY
done; done
