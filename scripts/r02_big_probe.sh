#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_engine_gpu.py -x -q -m gpu -k "big" 2>&1 | tail -15 > gpurun_out/big_tests.log
for n in 148 296; do
  timeout 600 python bench.py --workload C5 --ninit $n --steps 3 --warmup 3 --no-cpu-baseline --no-configs --no-parity > gpurun_out/big_c5_n${n}.json 2> gpurun_out/big_c5_n${n}.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/big_c5_n*.json")):
    try:
        d=json.load(open(f)); print(f, "%.3e steps/s"%d["value"], d["kernel_ms"], "iters/solve %.3f"%d["linsolve_iters_per_imr_solve"], "fp64 frac %.3f"%d["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e, open(f.replace('.json','.err')).read()[-500:])
PY
cat gpurun_out/big_tests.log
