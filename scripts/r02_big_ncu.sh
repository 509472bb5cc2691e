#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"bfwd|badj" -c 2 -o gpurun_out/r02_c5_probe -f python bench.py --workload C5 --ninit 148 --ntime 40 --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-parity > gpurun_out/ncu_big.log 2>&1
tail -3 gpurun_out/ncu_big.log
