#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "qubit or stored or stage or penalt" 2>&1 | tail -4
timeout 1500 python -m pytest tests/test_host_gpu.py tests/test_ref_regression_gpu.py -m gpu -x -q 2>&1 | tail -3
for ls in gmres neumann; do
  timeout 300 python bench.py --linsolve $ls --no-cpu-baseline --steps 3 > gpurun_out/bench_q_${ls}.json 2> gpurun_out/bench_q_${ls}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_q_${ls}.json"))
    print("$ls", "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], d["kernel_ms"], "fp64 %.3f" % d["fp64"]["frac"])
except Exception as ex:
    print("$ls failed", ex); print(open("gpurun_out/bench_q_${ls}.err").read()[-800:])
PY
done
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/sanitize_small.py 2>&1 | grep -E "RACECHECK|Race reported" | head -8
