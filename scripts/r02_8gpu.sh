#!/bin/bash
# scaling run as the driver launches it: N = 8 (and 4) GPUs, strong headline + weak companion, reference arm
mkdir -p gpurun_out
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
  echo "N=$n rc=$?"
done
python - <<'PY'
import json
for n in (8,4):
    f="gpurun_out/bench_%dgpu.json"%n
    try:
        d=json.load(open(f)); print(n, d["scaling"], "%.3e"%d["value"], "ms %.2f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], d["config"]["ninit_total"], d.get("weak"), d.get("kernel_ms"), d["clocks"])
    except Exception as e: print(f,"ERR",e, open(f.replace(".json",".err")).read()[-600:])
PY
