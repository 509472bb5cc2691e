#!/bin/bash
# two-GPU checks: entry-point tests that need 2 GPUs, strong/weak bench lines through the engine's own NCCL communicator
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_gpu.py -q -m gpu -k "two_gpus or indivisible" 2>&1 | tail -6 > gpurun_out/2gpu_tests.log
cat gpurun_out/2gpu_tests.log
timeout 600 python -X faulthandler -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_2gpu.json",):
    try:
        d=json.load(open(f)); print(f, d["scaling"], "%.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], d["config"]["ninit_total"], d.get("weak"), d.get("kernel_ms"))
    except Exception as e: print(f,"ERR",e)
PY
