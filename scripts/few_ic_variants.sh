mkdir -p gpurun_out
for E in 1 2 4; do
  QB_QUBIT_E=$E timeout 300 python bench.py --ninit 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_few32_E$E.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_few32_E$E.json')); print('E=$E ninit=32', '%.3e'%d['value'], d['kernel_ms'])"
done
