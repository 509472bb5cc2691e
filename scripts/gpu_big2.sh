#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_engine_gpu.py tests/test_ref_regression_gpu.py -m gpu -x -q -k "big or regression or compositional" 2>&1 | tail -6
for ls in gmres neumann; do
timeout 900 python bench.py --workload C5 --steps 2 --warmup 3 --linsolve $ls --no-cpu-baseline > gpurun_out/bench_c5_$ls.json 2>> gpurun_out/bench_c5.err
done
python - <<PY
import json
for n in ("gmres","neumann"):
    try:
        d=json.load(open(f"gpurun_out/bench_c5_{n}.json"))
        print(n, "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], d["kernel_ms"], "fp64 %.3f" % d["fp64"]["frac"], "rhs/s %.3e" % d["rhs_evals_per_s"])
    except Exception as ex: print(n, "failed", ex)
PY
