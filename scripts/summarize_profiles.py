#!/usr/bin/env python
"""Summarise gpurun_out/launches.csv (ncu launch list of a bench run) and gpurun_out/prof_sweep.ncu-rep (ncu --set full of the
sweep kernels) into profiles/<tag>_*.  Usage: python scripts/summarize_profiles.py r01"""
import csv
import io
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

# ---- launch list
path = os.path.join(ROOT, "gpurun_out", "launches.csv")
if os.path.exists(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v_ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
        name = r["Kernel Name"].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v_ms
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(out, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list, `bench.py --steps 2 --warmup 3 --no-cpu-baseline` ({tag})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares)\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {n} | {ms:.3f} | {100 * ms / tot:.2f} % |\n")
    print("wrote launches summary:", len(rows), "rows")

# ---- full capture
rep = os.path.join(ROOT, "gpurun_out", "prof_sweep.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    with open(os.path.join(out, f"{tag}_sweep_kernels_ncu.md"), "w") as f:
        f.write(f"# ncu --set full --clock-control none, sweep kernels ({tag}); workload in the file name of the command below\n\n")
        f.write("`ncu --set full --clock-control none --import-source on -k regex:sweep|qfwd|qadj -c 2 python bench.py --ntime 1000 --steps 1 --warmup 3`\n"
                "(C4 with 1000 instead of 10000 time steps so that the ~40 replays stay short; per-step behaviour is identical)\n\n")
        for r in rows[2:]:
            f.write("| metric | value | unit |\n|---|---|---|\n")
            for k in keys:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")
            f.write("\n")
    # stall reasons per kernel
    for pat in ("qfwd", "qadj", "fwd_sweep", "adj_sweep"):
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(src)))
        if len(rows) < 3:
            continue
        hdr = rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = defaultdict(int)
        total = 0
        for r in rows[2:]:
            try:
                total += int(r[ix["# Samples"]])
            except (ValueError, IndexError):
                continue
            for s in stalls:
                try:
                    tot[s] += int(r[ix[s]])
                except (ValueError, IndexError):
                    pass
        if total:
            with open(os.path.join(out, f"{tag}_sweep_kernels_ncu.md"), "a") as f:
                f.write(f"\n## warp stall sampling, `{rows[0][1][:80]}`\n\n| reason | share |\n|---|---|\n")
                for s, v in sorted(tot.items(), key=lambda x: -x[1])[:8]:
                    f.write(f"| {s} | {100 * v / total:.1f} % |\n")
    print("wrote full-capture summary")
