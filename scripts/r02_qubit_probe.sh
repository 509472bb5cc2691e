#!/bin/bash
# round-2 probe of the one-reduction qubit kernels: parity tests, then C4 throughput by elements-per-thread variant and batch size
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -m gpu -k "qubit or stored_stage or one_reduction or stages_only or forward_matches or adjoint_gradient" 2>&1 | tail -15 > gpurun_out/probe_tests.log
timeout 600 python -m pytest tests/test_host_gpu.py -x -q -m gpu -k "C4_gradient_gmres_full or C4_gradient_neumann or C3_gradient_random" 2>&1 | tail -15 >> gpurun_out/probe_tests.log
for E in ${PROBE_E:-4 2 1}; do
  for n in ${PROBE_N:-32 128 256 512}; do
    QB_QUBIT_E=$E timeout 300 python bench.py --ninit $n --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/probe_E${E}_n${n}.json 2> gpurun_out/probe_E${E}_n${n}.err
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/probe_E*_n*.json")):
    try:
        d=json.load(open(f)); print(f, "%.3e steps/s"%d["value"], d["kernel_ms"], "iters/solve %.3f"%d["linsolve_iters_per_imr_solve"])
    except Exception as e: print(f, "ERR", e)
PY
cat gpurun_out/probe_tests.log
if [ -n "$PROBE_NCU" ]; then
  QB_QUBIT_E=4 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"qfwd|qadj" -c 2 -o gpurun_out/r02_c4_probe -f python bench.py --ninit 256 --ntime 1000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_probe.log 2>&1
  ncu -i gpurun_out/r02_c4_probe.ncu-rep --page raw --csv > gpurun_out/r02_c4_probe_raw.csv 2>/dev/null
fi
