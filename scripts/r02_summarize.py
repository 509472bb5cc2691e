#!/usr/bin/env python
"""Turns the outputs of scripts/r02_final.sh (gpurun_out/) into the committed round-2 profile summaries (profiles/r02_*):
launch list, ncu --set full summaries of the three kernel families (with stall reasons and DRAM traffic against the algorithmic
bytes), the traffic file bench.py reads, the static SASS table, and copies of the bench lines."""
import csv
import io
import json
import os
import re
import shutil
import subprocess
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
OUT = os.path.join(ROOT, "profiles")
TAG = "r02"
os.makedirs(OUT, exist_ok=True)


def fnum(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


# ---- bench lines
for f in sorted(os.listdir(G)):
    if f.startswith("r02_bench_") and f.endswith(".json") and os.path.getsize(os.path.join(G, f)) > 10:
        shutil.copy(os.path.join(G, f), os.path.join(OUT, f))
for n in (2, 4, 8):
    p = os.path.join(G, f"bench_{n}gpu.json")
    if os.path.exists(p) and os.path.getsize(p) > 10:
        txt = [l for l in open(p) if l.startswith("{")]
        if txt:
            open(os.path.join(OUT, f"r02_bench_{n}gpu.json"), "w").write(txt[-1])

# ---- launch list
path = os.path.join(G, "launches.csv")
if os.path.exists(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = fnum(r["Metric Value"])
        unit = r.get("Metric Unit", "ns")
        v_ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
        agg[r["Kernel Name"].split("(")[0]][0] += 1
        agg[r["Kernel Name"].split("(")[0]][1] += v_ms
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, f"{TAG}_launches.md"), "w") as f:
        f.write(f"# ncu launch list of `bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-parity` ({TAG}, C4, 1 GPU)\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES with the live numbers of the bench line)\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {n} | {ms:.3f} | {100 * ms / tot:.2f} % |\n")
    print("launch list:", len(agg), "kernels")

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed_op_global_st.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum"]


def full_capture(rep, title, command, alg_bytes, note):
    """alg_bytes: dict kernel-substring -> algorithmic HBM bytes of that launch"""
    rawcsv = rep.replace(".ncu-rep", "_raw.csv")
    if os.path.exists(rawcsv):
        raw = open(rawcsv).read()
    elif os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        return {}
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = {}
    md = [f"# {title}\n", f"`{command}`\n", note + "\n"]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d["Kernel Name"]
        md.append(f"## `{name}`\n\n| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d and d[k] not in ("", "n/a"):
                md.append(f"| {k} | {d[k]} | {u[k]} |")
        st = sorted(((fnum(d[k]) or 0.0, k) for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")), reverse=True)
        md.append("\nwarp stall reasons (stalled warps per issue, top 7): " +
                  ", ".join(f"{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, k in st[:7]))
        tot_st = sum(v for v, _ in st) or 1.0
        glob = sum(v for v, k in st if any(t in k for t in ("long_scoreboard", "lg_throttle", "drain")))
        md.append(f"\nglobal-memory stalls (long_scoreboard + lg_throttle + drain): {glob:.3f} of {tot_st:.2f} stalled warps per issue = **{100 * glob / tot_st:.1f} %** of the warp time")
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        rd = (fnum(d.get("dram__bytes_read.sum", "")) or 0.0) * mult.get(u.get("dram__bytes_read.sum", "byte"), 1.0)
        wr = (fnum(d.get("dram__bytes_write.sum", "")) or 0.0) * mult.get(u.get("dram__bytes_write.sum", "byte"), 1.0)
        for sub, ab in alg_bytes.items():
            if sub in name:
                md.append(f"\nDRAM traffic {rd + wr:.4g} B (read {rd:.4g}, write {wr:.4g}) against {ab:.4g} algorithmic bytes of this launch: **{(rd + wr) / ab:.2f}x**")
                res[sub] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "algorithmic_bytes": ab, "kernel": name}
        md.append("")
    return res, "\n".join(md)


traffic = {}
# C4: 256 initial conditions x 1000 steps x 4 KB stages
c4 = full_capture(os.path.join(G, "prof_sweep.ncu-rep"), "ncu --set full --clock-control none: C4 sweep kernels (qubit-system kernels, round 2)",
                  "ncu --set full --clock-control none --import-source on -k regex:qfwd|qadj -c 2 python bench.py --ntime 1000 --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-parity",
                  {"qfwd": 256 * 1000 * 4096.0, "qadj": 256 * 1000 * 4096.0},
                  "C4 with 1000 instead of 10000 time steps so that ncu's replays stay short (per-step behaviour is identical); 256 initial conditions, "
                  "stages-only storage.  Algorithmic bytes: 16*dim B per step written by the forward sweep and read by the adjoint sweep (SURVEY 8d).")
if c4:
    open(os.path.join(OUT, f"{TAG}_sweep_kernels_ncu.md"), "w").write(c4[1])
    traffic["C4"] = {k2: dict(v, ninit=256, ntime=1000) for k2, v in (("fwd", c4[0].get("qfwd")), ("adj", c4[0].get("qadj"))) if v}
c5 = full_capture(os.path.join(G, "prof_big.ncu-rep"), "ncu --set full --clock-control none: C5 sweep kernels (large-state kernels, round 2)",
                  "ncu --set full --clock-control none --import-source on -k regex:bfwd|badj -c 2 python bench.py --workload C5 --ninit 148 --ntime 40 --steps 1 --warmup 3 ...",
                  {"bfwd": 148 * 40 * 65536.0, "badj": 148 * 40 * 65536.0},
                  "C5 (4,4,4), dim 4096, 148 initial conditions (one per SM), 40 steps.  Algorithmic bytes: 64 KB per step and initial condition (stage written / read once).")
if c5:
    open(os.path.join(OUT, f"{TAG}_big_kernels_ncu.md"), "w").write(c5[1])
    traffic["C5"] = {k2: dict(v, ninit=148, ntime=40) for k2, v in (("fwd", c5[0].get("bfwd")), ("adj", c5[0].get("badj"))) if v}
gen = []
for nm, nin, dimp in (("c1", 4, 32), ("c2", 16, 32)):
    g = full_capture(os.path.join(G, f"prof_generic_{nm}.ncu-rep"), f"ncu --set full: general small-N sweep kernels on {nm.upper()} (round 2)",
                     f"ncu --set full --clock-control none --import-source on -k regex:fwd_sweep|adj_sweep -c 2 python bench.py --workload {nm.upper()} --steps 1 --warmup 3 ...",
                     {"fwd_sweep": nin * 1000 * 16.0 * dimp, "adj_sweep": nin * 1000 * 16.0 * dimp},
                     f"{nm.upper()}: {nin} initial conditions, 1000 steps; one thread per element, no bulk-copy rings (trajectory through LDG/STG): see the global load/store instruction counts against the total.")
    if g:
        gen.append(g[1])
if gen:
    open(os.path.join(OUT, f"{TAG}_generic_kernels_ncu.md"), "w").write("\n\n".join(gen))
if traffic:
    json.dump(traffic, open(os.path.join(OUT, f"{TAG}_traffic.json"), "w"), indent=1)
    print("traffic:", {k: {kk: round((vv["dram_bytes_read"] + vv["dram_bytes_write"]) / vv["algorithmic_bytes"], 2) for kk, vv in v.items()} for k, v in traffic.items()})

# ---- static SASS mix
objs = [("qb_qubit_v1.o", "qfwd_kernelILi4ELb1ELi4ELb1E", "qq::qfwd_kernel<4,true,4,gmres>"), ("qb_qubit_v1.o", "qadj_kernelILi4ELb1ELi4ELb1ELb1E", "qq::qadj_kernel<4,true,4,stored,gmres>"),
        ("qb_qubit_v2.o", "qfwd_kernelILi4ELb1ELi2ELb1E", "qq::qfwd_kernel<4,true,2,gmres>"), ("qb_qubit_v2.o", "qadj_kernelILi4ELb1ELi2ELb1ELb1E", "qq::qadj_kernel<4,true,2,stored,gmres>"),
        ("qb_big_q3.o", "bfwd_kernelILi3ELb1ELb1ELi4E", "qbig::bfwd_kernel<3,true,true,4>"), ("qb_big_q3.o", "badj_kernelILi3ELb1ELb1ELi4E", "qbig::badj_kernel<3,true,true,4>"),
        ("qb_q2.o", "fwd_sweep_kernelILi2ELb1ELb0ELb1ELi32E", "qbk::fwd_sweep_kernel<2,all2,noJ,Lindblad,32> (C2)"), ("qb_q2.o", "adj_sweep_kernelILi2ELb1ELb0ELb0ELi32E", "qbk::adj_sweep_kernel<2,all2,noJ,Schroedinger,32> (C1)"),
        ("qb_dense.o", "dense_apply_kernel", "qdense::dense_apply_kernel")]
ops = ["DFMA", "DADD", "DMUL", "DMMA", "LDS", "STS", "SHFL", "BAR", "LDG", "STG", "LD", "ST", "LDL", "STL", "UBLKCP", "SYNCS", "MUFU"]
lines = ["# Static SASS instruction mix of the sweep kernels (cuobjdump -sass quandary_b200/build/*.o, sm_100a), round 2\n",
         "UBLKCP = cp.async.bulk (TMA bulk copies of stage / state trajectory and control-record rings), SYNCS = mbarrier operations, DMMA = fp64 tensor-core MMA; "
         "LDL/STL = local memory (spills or out-of-line argument passing).\n",
         "| kernel | instructions | " + " | ".join(ops) + " |", "|---|---|" + "---|" * len(ops)]
res_lines = []
for obj, sub, label in objs:
    p = os.path.join(ROOT, "quandary_b200", "build", obj)
    if not os.path.exists(p):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", p], capture_output=True, text=True).stdout.split("\n")
    cnt, tot, inside = Counter(), 0, False
    for l in sass:
        if "Function :" in l:
            inside = sub in l
            continue
        if inside:
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_]+)", l)
            if m:
                tot += 1
                cnt[m.group(2)] += 1
    if tot:
        lines.append(f"| {label} | {tot} | " + " | ".join(str(cnt.get(o, 0)) for o in ops) + " |")
    ru = subprocess.run(["cuobjdump", "--dump-resource-usage", p], capture_output=True, text=True).stdout.split("\n")
    for i, l in enumerate(ru):
        if "Function" in l and sub in l and i + 1 < len(ru):
            res_lines.append(f"* {label}: `{ru[i + 1].strip()}`")
lines.append("\nResource usage (cuobjdump --dump-resource-usage):\n")
lines += res_lines
open(os.path.join(OUT, f"{TAG}_sass_static.md"), "w").write("\n".join(lines) + "\n")
print("sass table:", len(res_lines), "kernels")

# ---- sanitizer
for tool in ("racecheck", "memcheck"):
    p = os.path.join(G, f"{tool}.log")
    if os.path.exists(p):
        keep = [l for l in open(p) if "SUMMARY" in l or l.startswith("C") or l.startswith("G") or "big" in l.lower()]
        open(os.path.join(OUT, f"{TAG}_sanitizer.txt"), "a" if tool == "memcheck" else "w").write(f"== compute-sanitizer --tool {tool} python scripts/sanitize_small.py ==\n" + "".join(keep[-40:]))
