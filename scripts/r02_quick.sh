#!/bin/bash
# quick A/B of a kernel change: the C4/C3 parity tests, then the headline bench at two batch sizes (no CPU arm, no extra configs)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "C4 or C3 or gmres or mono or spec or qubit" > gpurun_out/q_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/q_tests.log
for i in 1 2; do
  timeout 300 python bench.py --no-cpu-baseline --no-configs --no-parity > gpurun_out/q_bench_$i.json 2> gpurun_out/q_bench_$i.err
done
timeout 300 python bench.py --no-cpu-baseline --no-configs --no-parity --ninit 32 > gpurun_out/q_bench_32.json 2> gpurun_out/q_bench_32.err
timeout 300 python bench.py --workload C3 --no-cpu-baseline --no-configs --no-parity > gpurun_out/q_bench_cfg.json 2> gpurun_out/q_bench_cfg.err
tail -3 gpurun_out/q_tests.log
python - <<'PY'
import json
for f in ("q_bench_1","q_bench_2","q_bench_32","q_bench_cfg"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), {k:v.get("value") for k,v in (d.get("configs") or {}).items()} if isinstance(d.get("configs"),dict) else "")
    except Exception as e:
        print(f, "ERR", e)
PY
