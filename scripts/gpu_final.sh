#!/bin/bash
# end-of-round GPU session: full parity suite, smoke, headline bench lines (+ reference arm), ncu launch list and full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_default.err; cut -c1-300 gpurun_out/bench_reference.json
timeout 600 python bench.py --linsolve neumann --no-cpu-baseline > gpurun_out/bench_neumann.json 2>> gpurun_out/bench_default.err
timeout 600 python bench.py --stage recompute --no-cpu-baseline > gpurun_out/bench_recompute.json 2>> gpurun_out/bench_default.err
timeout 900 python bench.py --workload C5 --cpu-ntime 20 --steps 2 > gpurun_out/bench_c5.json 2>> gpurun_out/bench_default.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/sanitize_small.py > gpurun_out/racecheck.log 2>&1; grep -E "RACECHECK SUMMARY" gpurun_out/racecheck.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qfwd\|qadj -c 2 -o gpurun_out/prof_sweep -f python bench.py --ntime 1000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
python - <<PY
import json
for n in ("default","neumann","recompute","c5"):
    try:
        d=json.load(open(f"gpurun_out/bench_{n}.json"))
        print(n, "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], d["kernel_ms"], "fp64 %.3f" % d["fp64"]["frac"], d["clocks"])
    except Exception as ex: print(n, "failed", ex)
PY
