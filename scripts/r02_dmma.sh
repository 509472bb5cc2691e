#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py -x -q -m gpu -k "tensor_core or big_rhs" 2>&1 | tail -8 > gpurun_out/dmma_tests.log
timeout 600 python scripts/r02_dmma_ab.py > gpurun_out/dmma_ab.json 2> gpurun_out/dmma_ab.err
cat gpurun_out/dmma_tests.log; cat gpurun_out/dmma_ab.json; tail -3 gpurun_out/dmma_ab.err
timeout 600 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_fp64_op_dmma.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"dense_apply|bapply|dmma_peak" -c 12 --csv --log-file gpurun_out/dmma_ncu.csv python scripts/r02_dmma_ab.py > /dev/null 2>&1
tail -15 gpurun_out/dmma_ncu.csv | cut -c1-400
