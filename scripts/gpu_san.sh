#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_gpu.py -m gpu -q -k "spline0 or two_segments" 2>&1 | tail -3
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/sanitize_small.py > gpurun_out/racecheck.log 2>&1; tail -25 gpurun_out/racecheck.log
timeout 1200 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/memcheck.log 2>&1; tail -6 gpurun_out/memcheck.log
