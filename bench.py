#!/usr/bin/env python
"""bench.py -- Lindblad time-steps/sec (forward + adjoint, all initial density matrices) of the B200 engine, beside the
reference's own CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo's engine (N>1: launched by torchrun)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  # the unmodified reference on the host cores

Workload: BASELINE.json config 4 (SURVEY.md section 8d "C4"): 4 qubits, N=16, Lindblad (decay+dephasing), the full
`basis` of 256 initial density matrices, 10 000 implicit-midpoint steps, runtype=gradient (forward sweep with stored
trajectory + adjoint sweep with the reduced gradient), linear solves by GMRES with the reference's stopping rule.
One bench "step" = one optimizer evaluation = ninit x ntime fused (forward+adjoint) time steps.

value  = fused time-steps/s over all GPUs, states resident in HBM (control table rebuilt from a fresh design vector,
         forward sweep, adjoint sweep, gradient reduction, NCCL all-reduce of the gradient when N>1).
e2e    = the same evaluation through the host-buffer C ABI calls (qb_engine_set_params / qb_engine_forward /
         qb_engine_adjoint equivalents with pinned HOST buffers: H2D of alpha, rho0, rhoT_bar and D2H of rhoT, gradient
         inside the timed region).
Multi-GPU: BASELINE config 4 is "256 init rho ... init-cond sharded 8 GPU": the headline at N GPUs is that job -- the 256
initial conditions divided over the ranks as the reference's comm_init does (src/optimproblem.cpp:245-248), one NCCL all-reduce
of [gradient | 7 scalars] per evaluation through the engine's own C-ABI communicator (qb_engine_comm_init /
qb_engine_allreduce_grad_dev) -- i.e. "scaling": "strong".  The `weak` object of the JSON line gives the companion measurement
with 256 initial conditions PER GPU.
At N=1 the line also carries `configs` (short device-timed lines for BASELINE configs C1, C2, C3, C5 with their own CPU
numbers) and `parity_check` (the .cfg entry point run on the benchmarked configuration against the committed output files of
the reference binary).
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "Lindblad time-steps/sec (fwd+adj, all init rho)"
UNIT = "steps/s"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "quandary")


def flops_per_rhs(nlevels, lindblad, Jkl):
    """Nominal flop model of one application of the matrix-free operator."""
    Q = len(nlevels)
    N = 1
    for n in nlevels:
        N *= n
    if lindblad:
        per = 6.0
        for n in nlevels:
            per += 8.0 * (4.0 * (n - 1) / n) + 4.0 * ((n - 1) / n) ** 2
        idx = 0
        for k in range(Q):
            for l in range(k + 1, Q):
                if abs(Jkl[idx]) > 1e-10:
                    per += 8.0 * 4.0 * ((nlevels[k] - 1) / nlevels[k]) * ((nlevels[l] - 1) / nlevels[l])
                idx += 1
        return per * N * N
    per = 2.0
    for n in nlevels:
        per += 8.0 * (2.0 * (n - 1) / n)
    return per * N


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        if shutil.which("nvidia-smi") is None:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, reasons, smmax, power = [], set(), None, []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smmax = float(f[2]); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=smmax, reasons=sorted(reasons), samples=len(sm), power_w_max=max(power))
        return out


# ------------------------------------------------------------------------------------------------------------------
# CPU side: the unmodified reference (oracle/_ref/quandary = /root/reference/src compiled against the sequential
# PETSc/MPI shim, processes over initial conditions like `mpirun -np`), or the C port when that binary is absent.
# ------------------------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_initialcondition(ninit_full):
    """full `basis` for the small systems; the N `diagonal` states for C5 (its 4096-element basis would take minutes on a CPU)"""
    return "basis" if ninit_full <= 1024 else "diagonal"


def run_reference_once(workdir, name, ntime, linsolve, alpha, nproc, initialcondition="basis", runtype="gradient"):
    """One runtype=gradient evaluation of the reference binary; returns (seconds as timed by src/main.cpp:396,453, kind)."""
    from quandary_b200.systems import cfg_text

    with open(os.path.join(workdir, "params.dat"), "w") as f:
        f.write("\n".join(repr(float(x)) for x in alpha) + "\n")
    with open(os.path.join(workdir, "bench.cfg"), "w") as f:
        f.write(cfg_text(name, ntime=ntime, linsolve=linsolve, runtype=runtype, initialcondition=initialcondition))
    env = dict(os.environ)
    env["QSHIM_NP"] = str(nproc)
    r = subprocess.run([REF_BIN, "bench.cfg", "--quiet"], cwd=workdir, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference binary failed: " + r.stderr[-500:])
    with open(os.path.join(workdir, "data_out", "timing.dat")) as f:
        return float(f.read().split()[1])


def run_port_once(name, ntime, linsolve, alpha, nthreads):
    import numpy as np

    import qb_oracle
    from quandary_b200.systems import basis_states, cfg_system

    desc, ninit = cfg_system(name, ntime=ntime, linsolve=linsolve, flags=1)
    N = 1
    for k in range(desc.nosc):
        N *= desc.nlevels[k]
    rho0 = basis_states(N, desc.lindblad != 0)
    rbar = np.random.default_rng(0).standard_normal(rho0.shape) / ninit
    orc = qb_oracle.load()
    t0 = time.perf_counter()
    orc.batch_gradient(desc, alpha, rho0, rbar, nthreads=nthreads)
    return time.perf_counter() - t0


def cpu_sample(name, ninit, ntime_sample, linsolve, alpha, reps=1, runtype="gradient"):
    """Times the CPU path on a bounded sample (same system and initial conditions, ntime_sample steps).
    Returns (kind, processes, seconds per repetition, initial conditions propagated)."""
    cores = host_cores()
    ic = cpu_initialcondition(ninit)
    if ic == "diagonal":
        ninit = int(round(ninit ** 0.5))
    nproc = 1
    while nproc * 2 <= min(cores, ninit, 64) and ninit % (nproc * 2) == 0:     # the shim's fork-MPI holds at most 64 ranks
        nproc *= 2     # the reference requires ninit % np_init == 0 (src/main.cpp:150-153)
    times = []
    if os.path.exists(REF_BIN):
        kind = "reference"
        wd = tempfile.mkdtemp(prefix="qb_bench_ref_")
        try:
            for _ in range(reps):
                times.append(run_reference_once(wd, name, ntime_sample, linsolve, alpha, nproc, ic, runtype))
        finally:
            shutil.rmtree(wd, ignore_errors=True)
    else:
        kind = "port"
        for _ in range(reps):
            times.append(run_port_once(name, ntime_sample, linsolve, alpha, nproc))
    return kind, nproc, times, ninit



def main_reference(args):
    """The reference's own CPU implementation of the path (oracle/_ref/quandary: unmodified reference sources, PETSc shim) on the
    box's host cores, on the b200 arm's headline configuration (C4, all 256 initial conditions at every N)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np

    from quandary_b200 import desc_ndesign
    from quandary_b200.systems import cfg_system, random_alpha

    desc, ninit = cfg_system(args.workload, ntime=args.ntime, linsolve=args.linsolve)
    alpha = random_alpha(desc_ndesign(desc), amp=0.005)
    nt = cpu_sample_ntime(args, ninit, alpha)
    kind, nproc, times, nic = cpu_sample(args.workload, ninit, nt, args.linsolve, alpha, reps=args.steps)
    total = float(np.sum(times))
    value = nic * nt * len(times) / total
    sample = (f"{nic} initial conditions x {nt} of {args.ntime} IMR steps per timed step, runtype=gradient (fwd+adj), {args.linsolve}, "
              f"{nproc} processes over initial conditions (the reference's np_init) on {host_cores()} host cores, timed as src/main.cpp:396,453")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, ninit, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nproc, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def cpu_sample_ntime(args, ninit, alpha):
    """IMR steps of the bounded CPU sample: --cpu-ntime if given, else sized by one short probe run to about --cpu-seconds of CPU
    work per timed step (at most the full sweep)"""
    if args.cpu_ntime:
        return min(args.ntime, args.cpu_ntime)
    probe = 200 if ninit <= 1024 else 10          # C5: 4 KB -> 64 KB states, the probe must stay a probe
    t_probe = cpu_sample(args.workload, ninit, probe, args.linsolve, alpha, reps=1)[2][0]
    return max(probe, min(args.ntime, int(probe * args.cpu_seconds / max(t_probe, 1e-3))))


WORKLOAD_TEXT = {
    "C1": "C1: 2-qubit closed-system Schroedinger N=4 (dim 4), CNOT, 4 basis states, 1000 IMR steps x dt 0.1, fwd+adj",
    "C2": "C2: 2-qubit Lindblad N=4 (dim 16), CNOT, 16 basis initial conditions, 1000 IMR steps x dt 0.1, forward only (runtype=simulation)",
    "C3": "C3: 3-qubit Lindblad N=8 (dim 64), Toffoli (cqnot), 64 basis initial conditions, 2000 IMR steps x dt 0.05, fwd+adj",
    "C4": "C4: 4-qubit Lindblad N=16 (dim 256), 256 basis initial conditions, 10000 IMR steps x dt 0.01, runtype=gradient (fwd+adj)",
    "C5": "C5: 3 qudits x 4 levels, Lindblad N=64 (dim 4096), Jkl couplings, 1000 IMR steps x dt 0.01, runtype=gradient (fwd+adj)",
}


def workload_config(args, ninit_total, world):
    return {
        "workload": WORKLOAD_TEXT.get(args.workload, args.workload) + f", linearsolver={args.linsolve} maxiter 20 atol 1e-10",
        "ninit_total": ninit_total, "ninit_per_gpu": ninit_total // world, "ntime": args.ntime, "linearsolver": args.linsolve, "adjoint_stage": args.stage,
        "forward_store": {1: "states + stages", 2: "stages only (the adjoint sweep's input)"}[args.store],
        "parallelism": f"initial conditions sharded over {world} GPU(s) (reference rule ninit/np_init per rank), one NCCL all-reduce of [gradient | 7 scalars] per evaluation",
        "cache": "stage store per evaluation (ninit x ntime x 4 KB = 10.5 GB at 256 x 10000) >> 126 MB L2; no flush needed",
    }


class Workload:
    """One engine + device-resident inputs of a named configuration; evaluation() = one optimizer evaluation."""

    def __init__(self, name, ntime, linsolve, stage, store, device, ninit=None, gradient=True):
        import numpy as np

        import quandary_b200 as qb
        from quandary_b200.systems import basis_states, cfg_system, random_alpha

        self.qb, self.np = qb, np
        flags = {"recompute": qb.FLAG_RECOMPUTE_STAGE, "stored": qb.FLAG_STORE_STAGE, "average": 0}[stage]
        self.desc, self.ninit_full = cfg_system(name, ntime=ntime, linsolve=linsolve, flags=flags)
        self.desc.device = device
        self.name, self.store, self.gradient = name, store, gradient
        self.nd = qb.desc_ndesign(self.desc)
        self.dim = qb.desc_dim(self.desc)
        self.lind = self.desc.lindblad != 0
        self.N = int(round(np.sqrt(self.dim))) if self.lind else self.dim
        self.ninit = ninit or self.ninit_full
        basis = basis_states(self.N, self.lind)
        self.rho0 = basis[np.arange(self.ninit) % self.ninit_full]
        self.alphas = [random_alpha(self.nd, seed=1234 + i, amp=0.005) for i in range(4)]
        self.eng = qb.Engine(self.desc)
        self.nlev = [self.desc.nlevels[k] for k in range(self.desc.nosc)]
        self.F = flops_per_rhs(self.nlev, self.lind, [self.desc.Jkl[i] for i in range(self.desc.nosc * (self.desc.nosc - 1) // 2)])

    def set_rbar(self, seed, scale):
        self.rbar = self.np.random.default_rng(seed).standard_normal(self.rho0.shape) * scale

    def upload(self, n):
        self.eng.upload_states(self.rho0[:n], 0)
        if self.gradient:
            self.eng.upload_states(self.rbar[:n], 1)

    def evaluation(self, i, n, allreduce=False):
        self.eng.set_params(self.alphas[i % len(self.alphas)])
        store = (self.store if self.lind else 0) if self.gradient else 0
        self.eng.forward_dev(n, store=store)
        if self.gradient:
            self.eng.adjoint_dev(n)
            if allreduce:
                self.eng.allreduce_grad_dev()

    def close(self):
        self.eng.close()


def device_timed(wl, n, steps, warmup, stream, dev, world, dist, allreduce):
    """W untimed + K timed evaluations with device-resident inputs, CUDA events on the engine's stream, max over ranks."""
    import torch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.cuda.stream(stream):
        wl.upload(n)
        for i in range(warmup):
            wl.evaluation(i, n, allreduce)
        barrier()
        s0 = wl.eng.stats()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        ev0.record(stream)
        for i in range(steps):
            wl.evaluation(i, n, allreduce)
        ev1.record(stream)
        barrier()
        t1 = time.time()
        ms = ev0.elapsed_time(ev1)
        s1 = wl.eng.stats()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), s0, s1, t0, t1


def sweep_numbers(wl, n, s0, s1):
    """per-evaluation work counters and kernel times from the engine's statistics"""
    n_adj = max(1, s1["n_adj"] - s0["n_adj"])
    n_fwd = max(1, s1["n_fwd"] - s0["n_fwd"])
    fwd_ms = (s1["total_fwd_ms"] - s0["total_fwd_ms"]) / n_fwd
    adj_ms = (s1["total_adj_ms"] - s0["total_adj_ms"]) / n_adj if wl.gradient else 0.0
    solves = (2.0 if wl.gradient else 1.0) * n * wl.desc.ntime
    if wl.gradient and not wl.lind:
        solves = 3.0 * n * wl.desc.ntime        # Schroedinger adjoint: backward primal step + adjoint solve + (recomputed) stage
    return fwd_ms, adj_ms, s1["rhs_evals"], s1["linsolve_iters"] / solves


def parity_check():
    """Runs the .cfg entry point on the benchmarked configuration (C4, 10000 steps, GMRES, 256 initial conditions) and compares
    gradient, objective and fidelity with the files the unmodified reference binary wrote for it (tests/golden/host/C4_gradient_gmres_full,
    tests/golden/make_golden_host.py)."""
    import numpy as np

    gold = os.path.join(ROOT, "tests", "golden", "host", "C4_gradient_gmres_full")
    binary = os.path.join(ROOT, "quandary_b200", "quandary_b200_main")
    if not (os.path.exists(gold) and os.path.exists(binary)):
        return {"skipped": "golden case or entry-point binary missing"}

    def load(path):
        return np.array([[float(t) for t in ln.split()] for ln in open(path) if ln.strip() and not ln.startswith("#")])

    wd = tempfile.mkdtemp(prefix="qb_bench_parity_")
    try:
        for f in ("case.cfg", "params.dat"):
            shutil.copy(os.path.join(gold, f), wd)
        t0 = time.perf_counter()
        r = subprocess.run([binary, "case.cfg", "--quiet"], cwd=wd, capture_output=True, text=True, timeout=900)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            return {"failed": (r.stdout + r.stderr)[-300:]}
        g, gr = load(os.path.join(wd, "data_out", "grad.dat")), load(os.path.join(gold, "base", "grad.dat"))
        h, hr = load(os.path.join(wd, "data_out", "optim_history.dat")), load(os.path.join(gold, "base", "optim_history.dat"))
        return {"golden": "tests/golden/host/C4_gradient_gmres_full (output files of the unmodified reference binary, 10000 steps x 256 initial conditions, gmres)",
                "grad_rel_err": float(np.max(np.abs(g - gr)) / np.max(np.abs(gr))), "objective_rel_err": float(abs(h[0, 1] - hr[0, 1]) / abs(hr[0, 1])),
                "fidelity_abs_err": float(abs(h[0, 4] - hr[0, 4])), "tolerance": 1e-10, "entry_point_wall_s": wall}
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def config_line(name, args, stream, dev, fp64_peak):
    """short device-timed line of another BASELINE config (its own batch, its own CPU number)"""
    import torch

    gradient = name != "C2"                         # C2 is "forward-only" in BASELINE.json
    ninit = 296 if name == "C5" else None           # two initial conditions per SM; the full 4096-element basis needs batching (2 x 268 GB)
    wl = Workload(name, None, args.linsolve, args.stage, args.store, dev.index, ninit=ninit, gradient=gradient)
    wl.eng.set_stream(stream.cuda_stream)
    wl.set_rbar(7, 1.0 / wl.ninit)
    ms, s0, s1, _, _ = device_timed(wl, wl.ninit, 3, 3, stream, dev, 1, None, False)
    fwd_ms, adj_ms, rhs_evals, iters = sweep_numbers(wl, wl.ninit, s0, s1)
    steps = wl.ninit * wl.desc.ntime * 3
    tf = wl.F * rhs_evals / ((fwd_ms + adj_ms) * 1e-3) / 1e12
    out = {"workload": WORKLOAD_TEXT[name], "ninit": wl.ninit, "ntime": wl.desc.ntime, "value": steps / (ms * 1e-3), "unit": UNIT,
           "ms_per_evaluation": ms / 3, "kernel_ms": {"fwd_sweep": fwd_ms, "adj_sweep": adj_ms}, "linsolve_iters_per_imr_solve": iters,
           "rhs_evals_per_s": rhs_evals * 3 / (ms * 1e-3),
           "fp64": {"achieved_tflops": tf, "peak_tflops_measured_dfma": fp64_peak, "frac": tf / fp64_peak if fp64_peak > 0 else None, "flops_per_rhs_eval": wl.F}}
    alpha = wl.alphas[0]
    ninit_full, ntime = wl.ninit_full, wl.desc.ntime
    wl.close()
    if not args.no_cpu_baseline:
        nt = {"C1": ntime, "C3": min(ntime, 500), "C5": 100}.get(name, ntime)
        kind, nproc, times, nic = cpu_sample(name, ninit_full, nt, args.linsolve, alpha, reps=1, runtype="gradient" if gradient else "simulation")
        if kind == "port" and not gradient:
            times = [t / 2.0 for t in times]       # the port times forward + adjoint; forward-only is reported as half (stated in the sample)
        out["cpu_baseline"] = {"value": nic * nt / times[0], "unit": UNIT, "cores": nproc, "kind": kind,
                               "sample": f"{nic} initial conditions ({cpu_initialcondition(ninit_full)}) x {nt} of {ntime} steps, {'fwd+adj' if gradient else 'forward only'}, "
                                         f"{nproc} processes on {host_cores()} host cores ({times[0]:.1f} s)"}
    return out


# ------------------------------------------------------------------------------------------------------------------
def main_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import quandary_b200 as qb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # rank 0 prints exactly one JSON line: NCCL's version banner (printed to stdout at the first communicator of a process when
    # NCCL_DEBUG >= VERSION) is sent to stderr at the file-descriptor level while the communicators are created
    class _StdoutToStderr:
        def __enter__(self):
            sys.stdout.flush()
            self.saved = os.dup(1)
            os.dup2(2, 1)

        def __exit__(self, *a):
            sys.stdout.flush()
            os.dup2(self.saved, 1)
            os.close(self.saved)

    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        with _StdoutToStderr():
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()

    wl = Workload(args.workload, args.ntime, args.linsolve, args.stage, args.store, local_rank, ninit=args.ninit)
    ninit_full = wl.ninit if args.ninit else wl.ninit_full
    strong = world > 1 and args.scaling == "strong" and ninit_full % world == 0
    n_head = ninit_full // world if strong else ninit_full          # initial conditions of this rank in the headline measurement
    # rank r owns initial conditions [r*n_head, (r+1)*n_head) of the job (src/optimproblem.cpp:248)
    rho0_all = wl.rho0
    if strong:
        wl.rho0 = rho0_all[rank * n_head:(rank + 1) * n_head]
    total_head = n_head * world
    wl.set_rbar(1000 + rank, 1.0 / total_head)

    stream = torch.cuda.Stream(device=dev)
    wl.eng.set_stream(stream.cuda_stream)
    if world > 1:
        # the engine's own communicator (C ABI): rank 0 creates the NCCL id, torch.distributed only carries the 128 bytes
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(qb.comm_unique_id()), dtype=torch.uint8))
        with _StdoutToStderr():
            dist.broadcast(idt, 0)
            wl.eng.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
            wl.eng.allreduce_grad_dev()      # first collective of the communicator (connection set-up) outside the timed region
            torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, s0, s1, t0, t1 = device_timed(wl, n_head, args.steps, args.warmup, stream, dev, world, dist, world > 1)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ntime = wl.desc.ntime
    steps_total = total_head * ntime * args.steps
    value = steps_total / (ms * 1e-3)
    launches = s1["kernel_launches"] - s0["kernel_launches"]
    fwd_ms, adj_ms, rhs_evals, iters = sweep_numbers(wl, n_head, s0, s1)

    # ---- end to end through the host-buffer ABI (pinned host memory, copies inside the timed region) ----
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    h_rho0, h_rbar = pin(wl.rho0[:n_head]), pin(wl.rbar[:n_head])
    h_alpha = [pin(a) for a in wl.alphas]
    e2e_times = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.cuda.stream(stream):
        for i in range(args.warmup + args.steps):
            barrier()
            w0 = time.perf_counter()
            wl.eng.set_params(h_alpha[i % len(h_alpha)])
            wl.eng.upload_states(h_rho0, 0)
            wl.eng.forward_dev(n_head, store=args.store)
            h_rhoT = wl.eng.download_states(n_head, 0)
            wl.eng.upload_states(h_rbar, 1)
            wl.eng.adjoint_dev(n_head)
            if world > 1:
                wl.eng.allreduce_grad_dev()
            h_grad = wl.eng.download_grad()
            barrier()
            if i >= args.warmup:
                e2e_times.append(time.perf_counter() - w0)
    te = torch.tensor([sum(e2e_times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = steps_total / float(te.item())
    h2d = h_alpha[0].nbytes + h_rho0.nbytes + h_rbar.nbytes
    d2h = h_rhoT.nbytes + h_grad.nbytes

    # ---- companion: weak scaling (the full batch on every GPU) ----
    weak = None
    if strong:
        wl.rho0 = rho0_all
        nw = ninit_full
        wl.set_rbar(2000 + rank, 1.0 / (nw * world))
        ms_w, _, _, _, _ = device_timed(wl, nw, args.steps, 2, stream, dev, world, dist, True)
        weak = {"ninit_total": nw * world, "ninit_per_gpu": nw, "value": nw * world * ntime * args.steps / (ms_w * 1e-3), "unit": UNIT, "ms_per_step": ms_w / args.steps}

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        dimp = (1024 if wl.dim <= 1024 else 4096) if wl.dim > 256 else (256 if wl.dim > 64 else max(32, wl.dim))
        traj_bytes = 16.0 * dimp * n_head * ntime           # written once by the forward sweep (the stages), read once by the adjoint sweep
        adj_gbs = traj_bytes / (adj_ms * 1e-3) / 1e9
        fwd_gbs = traj_bytes * (2 if args.store == 1 else 1) / (fwd_ms * 1e-3) / 1e9
        # DRAM traffic of the dominant kernel: ncu --set full capture committed under profiles/ (static: ncu cannot run inside a timed bench),
        # scaled linearly to this launch's size
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            k = tr.get(args.workload, {}).get("adj")
            if k:
                traffic = (k["dram_bytes_read"] + k["dram_bytes_write"]) * (ntime / k["ntime"]) * (n_head / k["ninit"])
                traffic_src = "static_from_profile: profiles/r02_traffic.json (ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of the adjoint sweep kernel, scaled by ninit x ntime)"
        except (OSError, ValueError, KeyError):
            pass
        kname = {"C4": "qq::qadj_kernel<4,true,E,stored,gmres> (+ qq::qfwd_kernel)", "C5": "qbig::badj_kernel (+ bfwd_kernel)"}.get(args.workload, "adjoint sweep kernel")
        fp64_peak = qb.measure_fp64_peak(local_rank)
        F = wl.F
        fp64_tf = F * rhs_evals / ((fwd_ms + adj_ms) * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, total_head, world),
            "rhs_evals_per_s": rhs_evals * world * args.steps / (ms * 1e-3),
            "rhs_evals_per_evaluation_per_gpu": rhs_evals, "linsolve_iters_per_imr_solve": iters,
            "kernel_ms": {"fwd_sweep": fwd_ms, "adj_sweep": adj_ms},
            "roofline": {"bound": "fp64", "kernel": kname, "achieved": fp64_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp64_tf / fp64_peak if fp64_peak > 0 else None,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": "DFMA microbenchmark run inside this bench (qb_measure_fp64_peak); MEASURED_PEAKS.json carries no fp64 figure",
                         "flops_per_rhs_eval": F, "flops_per_launch_pair": F * rhs_evals,
                         "note": "nominal stencil flops (SURVEY 8d) of all operator applications of the two sweep launches / their device time. The path is fp64- and "
                                 "latency-bound (arithmetic intensity ~50 flop/B vs a ridge of ~5.6): 256 independent sweeps of 10000 dependent steps each"},
            "hbm": {"achieved": adj_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": adj_gbs / hbm_peak, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": traj_bytes, "fwd_sweep_achieved": fwd_gbs,
                    "note": "the fused sweeps keep the state on chip: the only per-step HBM traffic is the stage trajectory, 16*dim B per step, written once and read once"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * float(te.item()) / args.steps},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if weak:
            line["weak"] = weak
    alpha0, ntime_full, ninit_system = wl.alphas[0], ntime, wl.ninit_full
    wl.close()
    if world > 1:
        dist.barrier()
    if rank == 0 and world == 1:
        if not args.no_cpu_baseline:
            # the CPU arm always runs the system's own initial-condition set (basis, or diagonal when the basis has more than 1024 elements)
            nt = cpu_sample_ntime(args, ninit_system, alpha0)
            kind, nproc, times, nic = cpu_sample(args.workload, ninit_system, nt, args.linsolve, alpha0, reps=1)
            line["cpu_baseline"] = {
                "value": nic * nt / times[0], "unit": UNIT, "cores": nproc, "kind": kind,
                "sample": f"{nic} initial conditions x {nt} of {ntime_full} IMR steps, fwd+adj, {args.linsolve}, {nproc} processes "
                          f"over initial conditions on {host_cores()} host cores ({times[0]:.1f} s)"}
        if not args.no_parity and args.workload == "C4":
            line["parity_check"] = parity_check()
        if not args.no_configs and args.workload == "C4":
            line["configs"] = {}
            for name in ("C1", "C2", "C3", "C5"):
                try:
                    line["configs"][name] = config_line(name, args, stream, dev, line["roofline"]["peak"])
                except Exception as ex:      # a failing side line must not lose the headline
                    line["configs"][name] = {"error": str(ex)[:300]}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C4")
    ap.add_argument("--ntime", type=int, default=None, help="IMR steps per sweep (default: the workload's, 10000 for C4)")
    ap.add_argument("--ninit", type=int, default=None, help="initial conditions of the job (default: the full basis, 256 for C4)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N>1: 'strong' divides the job's initial conditions over the ranks (BASELINE config 4), 'weak' gives every rank the full batch")
    ap.add_argument("--linsolve", default="gmres", choices=["gmres", "neumann"])
    ap.add_argument("--stage", default="stored", choices=["stored", "recompute", "average"],
                    help="adjoint primal stage x_{n-1}+dt/2*k: 'stored' = kept by the forward sweep (default; the value the reference re-solves "
                         "for); 'recompute' = re-solved in the adjoint step exactly as ImplMidpoint::evolveBWD does (src/timestepper.cpp:674-680); "
                         "'average' = (x_{n-1}+x_n)/2 (equal up to rounding)")
    ap.add_argument("--store", type=int, default=2, choices=[1, 2],
                    help="what the forward sweep keeps: 2 = the adjoint sweep's input only (the stages; the reference's storeFWD memory), "
                         "1 = the state trajectory as well (needed by penalty terms and trajectory output)")
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="CPU work per timed step of the reference arm / cpu_baseline (sizes the bounded sample)")
    ap.add_argument("--cpu-ntime", type=int, default=0, help="fix the IMR steps of the CPU sample instead (10000 = the full-length sweep)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short lines of BASELINE configs C1, C2, C3, C5")
    ap.add_argument("--no-parity", action="store_true", help="skip the entry-point parity check of the benchmarked configuration")
    args = ap.parse_args()
    if args.ntime is None:
        from quandary_b200.systems import system_params
        args.ntime = system_params(args.workload)["nt"]
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return main_reference(args)
    return main_b200(args)


if __name__ == "__main__":
    sys.exit(main())
