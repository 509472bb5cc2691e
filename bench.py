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
Multi-GPU: every rank owns `--ninit` initial conditions (weak scaling: the job is N x 256 density matrices, the
reference's initial-condition parallelism, src/optimproblem.cpp:245-248); the `strong` object of the JSON line gives the
same measurement with the 256 initial conditions of C4 divided over the ranks.
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


def run_reference_once(workdir, name, ntime, linsolve, alpha, nproc, initialcondition="basis"):
    """One runtype=gradient evaluation of the reference binary; returns (seconds as timed by src/main.cpp:396,453, kind)."""
    from quandary_b200.systems import cfg_text

    with open(os.path.join(workdir, "params.dat"), "w") as f:
        f.write("\n".join(repr(float(x)) for x in alpha) + "\n")
    with open(os.path.join(workdir, "bench.cfg"), "w") as f:
        f.write(cfg_text(name, ntime=ntime, linsolve=linsolve, runtype="gradient", initialcondition=initialcondition))
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


def cpu_sample(name, ninit, ntime_sample, linsolve, alpha, reps=1):
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
                times.append(run_reference_once(wd, name, ntime_sample, linsolve, alpha, nproc, ic))
        finally:
            shutil.rmtree(wd, ignore_errors=True)
    else:
        kind = "port"
        for _ in range(reps):
            times.append(run_port_once(name, ntime_sample, linsolve, alpha, nproc))
    return kind, nproc, times, ninit


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np

    from quandary_b200 import desc_ndesign
    from quandary_b200.systems import cfg_system, random_alpha

    desc, ninit = cfg_system(args.workload, ntime=args.cpu_ntime, linsolve=args.linsolve)
    alpha = random_alpha(desc_ndesign(desc), amp=0.005)
    cpu_sample(args.workload, ninit, args.cpu_ntime, args.linsolve, alpha, reps=max(0, args.warmup and 1))
    kind, nproc, times, nic = cpu_sample(args.workload, ninit, args.cpu_ntime, args.linsolve, alpha, reps=args.steps)
    total = float(np.sum(times))
    value = nic * args.cpu_ntime * len(times) / total
    sample = (f"{nic} initial conditions x {args.cpu_ntime} of {args.ntime} IMR steps, runtype=gradient, {args.linsolve}, "
              f"{nproc} processes over initial conditions (reference's np_init), timed as src/main.cpp:396,453")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, args.ninit or ninit, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nproc, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, ninit_per_gpu, world):
    return {
        "workload": (f"{args.workload}: 4-qubit Lindblad N=16 (dim 256), basis initial conditions, {args.ntime} IMR steps x dt 0.01, "
                     f"runtype=gradient (fwd+adj), linearsolver={args.linsolve} maxiter 20 atol 1e-10") if args.workload == "C4" else
                    (f"{args.workload}: 3 qudits x 4 levels, Lindblad N=64 (dim 4096), Jkl couplings, basis initial conditions, {args.ntime} IMR steps x dt 0.01, "
                     f"runtype=gradient (fwd+adj), linearsolver={args.linsolve} maxiter 20 atol 1e-10") if args.workload == "C5" else args.workload,
        "ninit_per_gpu": ninit_per_gpu, "ninit_total": ninit_per_gpu * world, "ntime": args.ntime, "linearsolver": args.linsolve, "adjoint_stage": args.stage,
        "parallelism": f"initial conditions sharded over {world} GPU(s), one NCCL all-reduce of the gradient per evaluation",
        "cache": "trajectory store per evaluation (ninit x (ntime+1) x 4 KB = 10.5 GB at 256 x 10000) >> 126 MB L2; no flush needed",
    }


# ------------------------------------------------------------------------------------------------------------------
def main_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import quandary_b200 as qb
    from quandary_b200.systems import basis_states, cfg_system, random_alpha

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)

    flags = {"recompute": qb.FLAG_RECOMPUTE_STAGE, "stored": qb.FLAG_STORE_STAGE, "average": 0}[args.stage]
    desc, ninit_full = cfg_system(args.workload, ntime=args.ntime, linsolve=args.linsolve, flags=flags)
    desc.device = local_rank
    nd = qb.desc_ndesign(desc)
    dim = qb.desc_dim(desc)
    N = int(round(np.sqrt(dim))) if desc.lindblad != 0 else dim
    ninit = args.ninit or ninit_full
    if args.workload == "C5" and args.ninit is None:
        ninit = 296       # two initial conditions per SM-sized wave; the full 4096-element basis needs 2 x 268 GB of trajectory
    basis = basis_states(N, desc.lindblad != 0)
    rho0 = basis[np.arange(ninit) % ninit_full]
    del basis
    rng = np.random.default_rng(1000 + rank)
    rbar = rng.standard_normal(rho0.shape) / (ninit * world)
    alphas = [random_alpha(nd, seed=1234 + i, amp=0.005) for i in range(4)]

    stream = torch.cuda.Stream(device=dev)
    eng = qb.Engine(desc)
    eng.set_stream(stream.cuda_stream)

    class _GradView:      # zero-copy torch view of the engine's device gradient buffer [nd + 8]
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}

    gview = torch.as_tensor(_GradView(eng.grad_dev_ptr(), nd + 8), device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def evaluation(i, n):
        """one optimizer evaluation with device-resident states"""
        eng.set_params(alphas[i % len(alphas)])
        eng.forward_dev(n, store=True)
        eng.adjoint_dev(n)
        if world > 1:
            dist.all_reduce(gview)

    def timed(n, steps, warmup):
        with torch.cuda.stream(stream):
            eng.upload_states(rho0[:n], 0)
            eng.upload_states(rbar[:n], 1)
            for i in range(warmup):
                evaluation(i, n)
            barrier()
            s0 = eng.stats()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.time()
            ev0.record(stream)
            for i in range(steps):
                evaluation(i, n)
            ev1.record(stream)
            barrier()
            t1 = time.time()
            ms = ev0.elapsed_time(ev1)
            s1 = eng.stats()
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), s0, s1, t0, t1

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, s0, s1, t0, t1 = timed(ninit, args.steps, args.warmup)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    steps_total = ninit * world * desc.ntime * args.steps
    value = steps_total / (ms * 1e-3)
    launches = s1["kernel_launches"] - s0["kernel_launches"]
    # per-evaluation work counters (rhs_evals/linsolve_iters are those of the last forward+adjoint pair)
    rhs_evals = s1["rhs_evals"]
    iters = s1["linsolve_iters"]
    n_adj = s1["n_adj"] - s0["n_adj"]
    n_fwd = s1["n_fwd"] - s0["n_fwd"]
    adj_ms = (s1["total_adj_ms"] - s0["total_adj_ms"]) / max(1, n_adj)
    fwd_ms = (s1["total_fwd_ms"] - s0["total_fwd_ms"]) / max(1, n_fwd)

    # ---- end to end through the host-buffer ABI (pinned host memory, copies inside the timed region) ----
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    h_rho0, h_rbar = pin(rho0), pin(rbar)
    h_alpha = [pin(a) for a in alphas]
    e2e_times = []
    with torch.cuda.stream(stream):
        for i in range(args.warmup + args.steps):
            barrier()
            w0 = time.perf_counter()
            eng.set_params(h_alpha[i % len(h_alpha)])
            eng.upload_states(h_rho0, 0)
            eng.forward_dev(ninit, store=True)
            h_rhoT = eng.download_states(ninit, 0)
            eng.upload_states(h_rbar, 1)
            eng.adjoint_dev(ninit)
            if world > 1:
                dist.all_reduce(gview)
            h_grad = eng.download_grad()
            barrier()
            if i >= args.warmup:
                e2e_times.append(time.perf_counter() - w0)
    te = torch.tensor([sum(e2e_times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = steps_total / float(te.item())
    h2d = h_alpha[0].nbytes + h_rho0.nbytes + h_rbar.nbytes
    d2h = h_rhoT.nbytes + h_grad.nbytes

    # ---- strong-scaling companion: the 256 initial conditions of C4 divided over the ranks ----
    strong = None
    if world > 1 and ninit_full % world == 0 and ninit_full // world <= rho0.shape[0]:
        nloc = ninit_full // world
        ms_s, _, _, _, _ = timed(nloc, args.steps, 1)
        strong = {"ninit_total": ninit_full, "ninit_per_gpu": nloc, "value": ninit_full * desc.ntime * args.steps / (ms_s * 1e-3), "unit": UNIT,
                  "ms_per_step": ms_s / args.steps}

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        dimp = (1024 if dim <= 1024 else 4096) if dim > 256 else (256 if dim > 64 else max(32, dim))
        traj_bytes = 16.0 * dimp * ninit * (desc.ntime + 1)      # read once by the adjoint sweep; the forward sweep writes it (twice with stored stages)
        adj_gbs = traj_bytes / (adj_ms * 1e-3) / 1e9
        fwd_gbs = traj_bytes / (fwd_ms * 1e-3) / 1e9
        # DRAM traffic of the dominant kernel from the committed ncu --set full capture, scaled linearly to this launch's size
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if args.workload == "C4":
                k = tr["qadj_kernel"]
                traffic = (k["dram_bytes_read"] + k["dram_bytes_write"]) * (desc.ntime / tr["ntime"]) * (ninit / tr["ninit"])
        except (OSError, ValueError, KeyError):
            pass
        kname = {"C4": "qq::qadj_kernel<4,true,4>", "C5": "qbig::badj_kernel<3,true,true,4>"}.get(args.workload, "adjoint sweep kernel")
        fp64_peak = qb.measure_fp64_peak(local_rank)
        F = flops_per_rhs([desc.nlevels[k] for k in range(desc.nosc)], desc.lindblad != 0, [desc.Jkl[i] for i in range(desc.nosc * (desc.nosc - 1) // 2)])
        fp64_tf = F * rhs_evals / ((fwd_ms + adj_ms) * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, ninit, world),
            "rhs_evals_per_s": rhs_evals * world * args.steps / (ms * 1e-3),
            "rhs_evals_per_evaluation_per_gpu": rhs_evals, "linsolve_iters_per_imr_solve": iters / ((3.0 if args.stage == "recompute" else 2.0) * ninit * desc.ntime),
            "kernel_ms": {"fwd_sweep": fwd_ms, "adj_sweep": adj_ms},
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": adj_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": adj_gbs / hbm_peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": traj_bytes, "fwd_sweep_achieved": fwd_gbs,
                         "note": "fused sweep keeps the state on chip: the only per-step HBM traffic is the trajectory (16*dim B per step, read once by "
                                 "the adjoint sweep; ncu dram bytes = algorithmic bytes, no re-reads); the kernel is bound by the fp64 dependency chain of "
                                 "256 independent sweeps on 148 SMs, see fp64"},
            "fp64": {"achieved_tflops": fp64_tf, "peak_tflops_measured_dfma": fp64_peak, "frac": fp64_tf / fp64_peak if fp64_peak > 0 else None,
                     "flops_per_rhs_eval": F},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * float(te.item()) / args.steps},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if strong:
            line["strong"] = strong
        if world == 1 and not args.no_cpu_baseline:
            desc_c, _ = cfg_system(args.workload, ntime=args.cpu_ntime, linsolve=args.linsolve)
            kind, nproc, times, nic = cpu_sample(args.workload, ninit_full, args.cpu_ntime, args.linsolve, alphas[0], reps=1)
            line["cpu_baseline"] = {
                "value": nic * args.cpu_ntime / times[0], "unit": UNIT, "cores": nproc, "kind": kind,
                "sample": f"{nic} initial conditions x {args.cpu_ntime} of {args.ntime} IMR steps, fwd+adj, {args.linsolve}, {nproc} processes "
                          f"over initial conditions on {host_cores()} host cores ({times[0]:.1f} s)"}
    eng.close()
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
    ap.add_argument("--ninit", type=int, default=None, help="initial conditions per GPU (default: the full basis, 256 for C4)")
    ap.add_argument("--linsolve", default="gmres", choices=["gmres", "neumann"])
    ap.add_argument("--stage", default="stored", choices=["stored", "recompute", "average"],
                    help="adjoint primal stage x_{n-1}+dt/2*k: 'stored' = kept by the forward sweep in a second trajectory buffer (default; "
                         "bitwise the value the reference re-solves for); 'recompute' = re-solved in the adjoint step exactly as "
                         "ImplMidpoint::evolveBWD does (src/timestepper.cpp:674-680); 'average' = (x_{n-1}+x_n)/2 (equal up to rounding)")
    ap.add_argument("--cpu-ntime", type=int, default=400, help="IMR steps of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
