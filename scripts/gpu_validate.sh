#!/bin/bash
# validation after changes that do not touch the C4/C5 sweep kernels: full GPU parity suite, smoke, headline bench line + reference arm, C5 line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cut -c1-400 gpurun_out/bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_default.err; cut -c1-300 gpurun_out/bench_reference.json
timeout 900 python bench.py --workload C5 --cpu-ntime 20 --steps 2 --no-cpu-baseline > gpurun_out/bench_c5_check.json 2>> gpurun_out/bench_default.err; cut -c1-300 gpurun_out/bench_c5_check.json
