#!/bin/bash
# C5 kernels after a change of the Krylov scratch footprint: parity tests of the large-state kernels, DRAM bytes / L2 hit rate of
# the two sweep kernels with every SM busy (and with a quarter of them), then the C5 bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "C5 or big or large or S888 or S2020 or qudit" > gpurun_out/fp_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/fp_tests.log
tail -2 gpurun_out/fp_tests.log
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct
B="python bench.py --workload C5 --ntime 40 --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-parity"
run() { timeout 400 ncu --metrics $M --clock-control none -k regex:"bfwd|badj" -c 2 --csv --log-file gpurun_out/fp_$1.csv $B --ninit $2 > gpurun_out/fp_$1.log 2>&1; }
run n148 148
run n37 37
timeout 600 python bench.py --workload C5 --ninit 296 --steps 3 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/fp_bench.json 2> gpurun_out/fp_bench.err
python - <<'PY'
import csv, json
for t in ("n148","n37"):
    try:
        rows=[r for r in csv.reader(open(f"gpurun_out/fp_{t}.csv")) if len(r)>10 and r[0].isdigit()]
        for r in rows: print(t, r[4][:24], r[-3], r[-1], r[-2])
    except Exception as e: print(t,"ERR",e)
try:
    d=json.loads(open("gpurun_out/fp_bench.json").read().strip().splitlines()[-1]); print("C5 bench",d["value"],d["ms_per_step"],d.get("kernel_ms"),d.get("parity_check"))
except Exception as e: print("bench ERR",e); print(open("gpurun_out/fp_bench.err").read()[-1500:])
PY
