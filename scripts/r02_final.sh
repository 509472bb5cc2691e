#!/bin/bash
# round-2 GPU session: full parity suite, smoke, bench lines (+ reference arm), ncu launch list, sanitizer, ncu --set full captures
# of the three kernel families.  Outputs under gpurun_out/; scripts/r02_summarize.py turns them into profiles/r02_*.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; cut -c1-400 gpurun_out/r02_bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench_default.err; cut -c1-300 gpurun_out/r02_bench_reference.json
timeout 600 python bench.py --linsolve neumann --no-cpu-baseline --no-configs --no-parity > gpurun_out/r02_bench_neumann.json 2>> gpurun_out/r02_bench_default.err
timeout 600 python bench.py --ninit 512 --no-cpu-baseline --no-configs --no-parity > gpurun_out/r02_bench_n512.json 2>> gpurun_out/r02_bench_default.err
timeout 600 python bench.py --store 1 --no-cpu-baseline --no-configs --no-parity > gpurun_out/r02_bench_store_states.json 2>> gpurun_out/r02_bench_default.err
timeout 900 python bench.py --workload C5 --ninit 296 --steps 3 --no-configs --no-parity > gpurun_out/r02_bench_c5.json 2>> gpurun_out/r02_bench_default.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-parity > gpurun_out/ncu_launches.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/sanitize_small.py > gpurun_out/racecheck.log 2>&1; grep -E "RACECHECK SUMMARY" gpurun_out/racecheck.log
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/memcheck.log 2>&1; grep -E "ERROR SUMMARY" gpurun_out/memcheck.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"qfwd|qadj" -c 2 -o gpurun_out/prof_sweep -f python bench.py --ntime 1000 --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-parity > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bfwd|badj" -c 2 -o gpurun_out/prof_big -f python bench.py --workload C5 --ninit 148 --ntime 40 --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-parity > gpurun_out/ncu_big.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fwd_sweep|adj_sweep" -c 2 -o gpurun_out/prof_generic_c1 -f python bench.py --workload C1 --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-parity > gpurun_out/ncu_generic.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fwd_sweep|adj_sweep" -c 2 -o gpurun_out/prof_generic_c2 -f python bench.py --workload C2 --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-parity >> gpurun_out/ncu_generic.log 2>&1
# gpurun brings back at most 64 MiB: keep the raw-metric and source pages as (compressed) csv, drop the reports
for r in prof_sweep prof_big prof_generic_c1 prof_generic_c2; do
  if [ -f gpurun_out/$r.ncu-rep ]; then
    ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null
    ncu -i gpurun_out/$r.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${r}_src.csv.gz
    rm -f gpurun_out/$r.ncu-rep
  fi
done
du -sh gpurun_out
python - <<PY
import json
for n in ("default","neumann","n512","store_states","c5"):
    try:
        d=json.load(open(f"gpurun_out/r02_bench_{n}.json"))
        print(n, "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], d["kernel_ms"], "fp64 %.3f" % d["roofline"]["frac"], d["clocks"], d.get("parity_check"))
    except Exception as ex: print(n, "failed", ex)
PY
