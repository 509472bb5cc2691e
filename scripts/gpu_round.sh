#!/bin/bash
# One GPU-box session: parity tests, smoke, bench lines, ncu launch list and one full capture of the sweep kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_gmres.json 2> gpurun_out/bench_gmres.err; cat gpurun_out/bench_gmres.json; tail -3 gpurun_out/bench_gmres.err
timeout 300 python bench.py --stage average --no-cpu-baseline > gpurun_out/bench_gmres_avg.json 2>> gpurun_out/bench_gmres.err; cat gpurun_out/bench_gmres_avg.json
timeout 300 python bench.py --linsolve neumann --no-cpu-baseline > gpurun_out/bench_neumann.json 2>> gpurun_out/bench_gmres.err; cat gpurun_out/bench_neumann.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_gmres.err; cat gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep -c 2 -o gpurun_out/prof_sweep python bench.py --ntime 1000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
