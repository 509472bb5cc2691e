#!/bin/bash
# quick GPU session: parity tests + short bench lines for each kernel variant
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for E in 4 8; do
  for ls in gmres neumann; do
    QB_QUBIT_E=$E timeout 300 python bench.py --linsolve $ls --no-cpu-baseline --steps 3 > gpurun_out/bench_E${E}_${ls}.json 2> gpurun_out/bench_E${E}_${ls}.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_E${E}_${ls}.json"))
    print("E=$E $ls", "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], d["kernel_ms"], "fp64 %.3f" % d["fp64"]["frac"], d["clocks"])
except Exception as ex:
    print("E=$E $ls failed", ex); print(open("gpurun_out/bench_E${E}_${ls}.err").read()[-800:])
PY
  done
done
