#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload C5 --steps 2 --warmup 3 --cpu-ntime 20 > gpurun_out/bench_c5_gmres.json 2> gpurun_out/bench_c5.err; tail -2 gpurun_out/bench_c5.err
timeout 900 python bench.py --workload C5 --steps 2 --warmup 3 --linsolve neumann --no-cpu-baseline > gpurun_out/bench_c5_neumann.json 2>> gpurun_out/bench_c5.err
python - <<PY
import json
for n in ("gmres","neumann"):
    try:
        d=json.load(open(f"gpurun_out/bench_c5_{n}.json"))
        print(n, "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], d["kernel_ms"], "fp64 %.3f" % d["fp64"]["frac"], "rhs/s %.3e" % d["rhs_evals_per_s"], d.get("cpu_baseline"))
    except Exception as ex: print(n, "failed", ex)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bfwd\|badj -c 2 -o gpurun_out/prof_big -f python bench.py --workload C5 --ntime 40 --ninit 148 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_big.log 2>&1
tail -2 gpurun_out/ncu_big.log | cut -c1-200
