#!/bin/bash
# variants of elements-per-thread for the 4-qubit kernels + ncu capture of the default
mkdir -p gpurun_out
for E in 4; do
  for ls in gmres neumann; do
    QB_QUBIT_E=$E timeout 300 python bench.py --linsolve $ls --no-cpu-baseline --steps 3 > gpurun_out/bench_E${E}_${ls}.json 2> gpurun_out/bench_E${E}_${ls}.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_E${E}_${ls}.json"))
    print("E=$E $ls", "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], d["kernel_ms"], "fp64 %.3f" % d["fp64"]["frac"])
except Exception as ex:
    print("E=$E $ls failed", ex); print(open("gpurun_out/bench_E${E}_${ls}.err").read()[-800:])
PY
  done
done
QB_QUBIT_E=${PROF_E:-4} timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep\|qfwd\|qadj -c 2 -o gpurun_out/prof_sweep -f python bench.py --ntime 1000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-200
