#!/bin/bash
# C5 forward sweep at full occupancy: what the stage store costs (QB_EXP=0x100 default cache policy, 0x200 no store at all)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "C5 or big" > gpurun_out/fp2_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/fp2_tests.log
tail -2 gpurun_out/fp2_tests.log
for x in 0 0x100 0x200; do
  QB_EXP=$x timeout 400 python bench.py --workload C5 --ninit 296 --steps 2 --warmup 2 --no-cpu-baseline --no-configs --no-parity > gpurun_out/fp2_bench_$x.json 2> gpurun_out/fp2_bench_$x.err
done
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct
B="python bench.py --workload C5 --ntime 40 --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-parity"
timeout 400 ncu --metrics $M --clock-control none -k regex:"bfwd|badj" -c 2 --csv --log-file gpurun_out/fp2_n148.csv $B --ninit 148 > gpurun_out/fp2_n148.log 2>&1
python - <<'PY'
import csv, json
for x in ("0","0x100","0x200"):
    try:
        d=json.loads(open(f"gpurun_out/fp2_bench_{x}.json").read().strip().splitlines()[-1]); print("QB_EXP",x,d["value"],d["ms_per_step"],d.get("kernel_ms"))
    except Exception as e: print(x,"ERR",e)
rows=[r for r in csv.reader(open("gpurun_out/fp2_n148.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows: print("n148", r[4][:24], r[-3], r[-1], r[-2])
PY
