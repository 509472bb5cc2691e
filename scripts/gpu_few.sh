#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for n in 32 128 256; do
  for E in 4 default; do
    if [ "$E" = "default" ]; then unset QB_QUBIT_E; else export QB_QUBIT_E=$E; fi
    timeout 300 python bench.py --ninit $n --no-cpu-baseline --steps 3 > gpurun_out/bench_n${n}_${E}.json 2> gpurun_out/bench_n${n}_${E}.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_n${n}_${E}.json"))
    print("ninit=$n E=$E", "value %.3e" % d["value"], d["kernel_ms"])
except Exception as ex:
    print("ninit=$n E=$E failed", ex); print(open("gpurun_out/bench_n${n}_${E}.err").read()[-600:])
PY
  done
done
