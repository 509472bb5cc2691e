#!/bin/bash
# steady-state wall time of one gradient evaluation through the .cfg entry point on C4 (256 initial conditions x 10000 steps)
# with the final-time objective on the device (default) and on the host (QB_HOST_OBJECTIVE=1)
mkdir -p gpurun_out; : > gpurun_out/objective_ab.txt
for mode in 0 1 1 0 0 1; do
  d=$(mktemp -d); cp scripts/c4_gradient.cfg $d/case.cfg
  (cd $d && QB_REPEAT_EVAL=20 QB_HOST_OBJECTIVE=$mode /root/repo/quandary_b200/quandary_b200_main case.cfg --quiet > run.log 2>&1)
  echo "host_objective=$mode $(grep 'Evaluation time' $d/run.log)" | tee -a gpurun_out/objective_ab.txt
done
