#!/bin/bash
# Where do the C5 kernels' DRAM writes come from?  (a) a quarter of the SMs busy (scratch footprint 14 MB << L2), (b) the same with the
# Krylov scratch marked L2-persisting, (c) all SMs with the persisting window; then the speed with and without the window.
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct
B="python bench.py --workload C5 --ntime 40 --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-parity"
run() { # tag, ninit, persist
  QB_L2_PERSIST=$3 timeout 400 ncu --metrics $M --clock-control none -k regex:"bfwd|badj" -c 2 --csv --log-file gpurun_out/l2p_$1.csv $B --ninit $2 > gpurun_out/l2p_$1.log 2>&1
}
run n37_off 37 0
run n37_on 37 1
run n148_off 148 0
run n148_on 148 1
for p in 0 1; do
  QB_L2_PERSIST=$p timeout 400 python bench.py --workload C5 --steps 3 --warmup 3 --no-cpu-baseline --no-configs --no-parity > gpurun_out/l2p_bench_$p.json 2> gpurun_out/l2p_bench_$p.err
done
grep -h QB_L2_PERSIST gpurun_out/l2p_*.log gpurun_out/l2p_bench_1.err | sort | uniq
python - <<'PY'
import csv, json
for t in ("n37_off","n37_on","n148_off","n148_on"):
    try:
        rows=[r for r in csv.reader(open(f"gpurun_out/l2p_{t}.csv")) if len(r)>10 and r[0].isdigit()]
        for r in rows: print(t, r[4][:24], r[-3], r[-1], r[-2])
    except Exception as e: print(t,"ERR",e)
for p in (0,1):
    try:
        d=json.loads(open(f"gpurun_out/l2p_bench_{p}.json").read().strip().splitlines()[-1]); print("persist",p,d["value"],d["ms_per_step"])
    except Exception as e: print(p,"ERR",e)
PY
