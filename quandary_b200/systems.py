"""Workload definitions: BASELINE.json configs C1-C5 plus generic-level parity cases.

`cfg_system(name)` builds the engine descriptor (angular units, as the reference holds them after load);
`cfg_text(name)` writes the SAME system as a reference `.cfg` (GHz units, as the reference parses them) so that the
unmodified reference (oracle/_ref) and the engine are driven by identical physics; the design vector is fed to the
reference through `control_initialization0 = file, <params>` (src/optimproblem.cpp:168-175).
"""
import math

import numpy as np

from . import make_desc

TWO_PI = 2.0 * math.pi


def system_params(name, ntime=None):
    """Plain parameters (GHz / ns, as in a .cfg) of the named system."""
    if name == "C1":  # tests/regression/cnot/cnot.cfg physics: (2,2) Schroedinger
        p = dict(nlev=[2, 2], lind="none", nt=ntime or 1000, dt=0.1,
                 trans=[4.10595, 4.81526], rot=[4.10595, 4.81526], selfk=[0.2198, 0.2252], crossk=[0.01], J=[0.0],
                 car=[[0.0, -0.01, 0.02], [0.0, -0.015, 0.01]], nspl=10, T1=[0, 0], T2=[0, 0], gate="cnot")
    elif name == "C2":
        p = dict(nlev=[2, 2], lind="both", nt=ntime or 1000, dt=0.1,
                 trans=[4.10595, 4.81526], rot=[4.10595, 4.81526], selfk=[0.2198, 0.2252], crossk=[0.01], J=[0.0],
                 car=[[0.0, -0.01, 0.02], [0.0, -0.015, 0.01]], nspl=10, T1=[56.0, 60.0], T2=[28.0, 30.0], gate="cnot")
    elif name == "C3":
        p = dict(nlev=[2, 2, 2], lind="both", nt=ntime or 2000, dt=0.05,
                 trans=[4.10595, 4.81526, 5.2], rot=[4.10595, 4.81526, 5.2], selfk=[0.2198, 0.2252, 0.21], crossk=[0.1, 0.01, 0.1],
                 J=[0.0] * 3, car=[[0.0, 0.05], [0.0, -0.05], [0.0, 0.03]], nspl=50, T1=[80.0, 90.0, 100.0], T2=[40.0, 45.0, 50.0],
                 gate="cqnot")
    elif name == "C4":
        p = dict(nlev=[2, 2, 2, 2], lind="both", nt=ntime or 10000, dt=0.01,
                 trans=[4.1, 4.2, 4.3, 4.4], rot=[4.1, 4.2, 4.3, 4.4], selfk=[0.2] * 4, crossk=[0.001] * 6, J=[0.0] * 6,
                 car=[[0.0, 0.01, -0.01]] * 4, nspl=15, T1=[100.0, 110.0, 120.0, 130.0], T2=[50.0, 55.0, 60.0, 65.0], gate="cqnot")
    elif name == "C5":  # 3 qudits x 4 levels, N = 64, dim = 4096, Jaynes-Cummings couplings
        p = dict(nlev=[4, 4, 4], lind="both", nt=ntime or 1000, dt=0.01,
                 trans=[4.1, 4.2, 4.3], rot=[4.1, 4.2, 4.3], selfk=[0.2] * 3, crossk=[0.001] * 3, J=[0.001] * 3,
                 car=[[0.0, -0.2], [0.0, -0.2], [0.0, -0.2]], nspl=15, T1=[100.0, 110.0, 120.0], T2=[50.0, 55.0, 60.0], gate=None)
    elif name == "G32J":  # generic levels with Jaynes-Cummings coupling, Lindblad
        p = dict(nlev=[3, 2], lind="both", nt=ntime or 200, dt=0.1,
                 trans=[4.1, 4.8], rot=[4.0, 4.6], selfk=[0.2, 0.22], crossk=[0.01], J=[0.005],
                 car=[[0.0, 0.2], [0.0]], nspl=8, T1=[50.0, 60.0], T2=[0.0, 30.0], gate=None)
    elif name == "G44J":  # two qudits, dim 256, coupling, decay only
        p = dict(nlev=[4, 4], lind="decay", nt=ntime or 100, dt=0.1,
                 trans=[4.1, 4.3], rot=[4.05, 4.2], selfk=[0.2, 0.21], crossk=[0.002], J=[0.003],
                 car=[[0.0, -0.2], [0.0, 0.21]], nspl=6, T1=[70.0, 80.0], T2=[0.0, 0.0], gate=None)
    elif name == "G232S":  # Schroedinger, mixed levels, coupling
        p = dict(nlev=[2, 3, 2], lind="none", nt=ntime or 200, dt=0.1,
                 trans=[4.1, 4.3, 4.6], rot=[4.1, 4.25, 4.5], selfk=[0.2, 0.21, 0.22], crossk=[0.002, 0.001, 0.003],
                 J=[0.004, 0.0, 0.002], car=[[0.0], [0.0, 0.1], [0.0]], nspl=7, T1=[0] * 3, T2=[0] * 3, gate=None)
    elif name == "G5":  # one oscillator, 5 levels, dephasing only
        p = dict(nlev=[5], lind="dephase", nt=ntime or 100, dt=0.1,
                 trans=[4.1], rot=[4.0], selfk=[0.2], crossk=[], J=[],
                 car=[[0.0, -0.2, -0.4]], nspl=9, T1=[0.0], T2=[20.0], gate=None)
    elif name == "Q5":  # five qubits, Schroedinger (N = 32)
        p = dict(nlev=[2] * 5, lind="none", nt=ntime or 100, dt=0.1,
                 trans=[4.1, 4.2, 4.3, 4.4, 4.5], rot=[4.1, 4.2, 4.3, 4.4, 4.45], selfk=[0.2] * 5,
                 crossk=[0.001 * (i + 1) for i in range(10)], J=[0.0, 0.002, 0.0, 0.0, 0.001, 0.0, 0.0, 0.0, 0.003, 0.0],
                 car=[[0.0]] * 5, nspl=5, T1=[0] * 5, T2=[0] * 5, gate=None)
    elif name == "G3":  # one 3-level oscillator (the reference's pipulse regression system)
        p = dict(nlev=[3], lind="both", nt=ntime or 200, dt=0.01,
                 trans=[4.10595], rot=[4.10595], selfk=[0.2198], crossk=[], J=[],
                 car=[[0.0, -0.2198]], nspl=12, T1=[20.0], T2=[40.0], gate=None)
    elif name == "G33":  # two qutrits, Lindblad, dim 81 (a level tuple the reference's matrix-free operator is instantiated for)
        p = dict(nlev=[3, 3], lind="both", nt=ntime or 100, dt=0.05,
                 trans=[4.1, 4.3], rot=[4.05, 4.2], selfk=[0.2, 0.21], crossk=[0.002], J=[0.003],
                 car=[[0.0, -0.2], [0.0, -0.21]], nspl=8, T1=[70.0, 80.0], T2=[30.0, 0.0], gate=None)
    elif name == "A320":  # the reference's "AxC" regression system: 3-level qubit x 20-level cavity, dim 3600
        p = dict(nlev=[3, 20], lind="both", nt=ntime or 100, dt=0.01,
                 trans=[4.41666, 6.84081], rot=[4.41666, 6.84081], selfk=[0.230, 0.0], crossk=[0.001176], J=[0.0],
                 car=[[0.0, -0.23], [0.0]], nspl=10, T1=[0.0, 31.7], T2=[0.0, 0.0], gate=None)
    elif name == "S2020":  # two 20-level oscillators, Schroedinger, N = 400 (large-state kernels; the reference has this tuple matrix-free)
        p = dict(nlev=[20, 20], lind="none", nt=ntime or 100, dt=0.02,
                 trans=[4.1, 4.3], rot=[4.08, 4.27], selfk=[0.004, 0.003], crossk=[0.0005], J=[0.002],
                 car=[[0.0, -0.05], [0.0, 0.04]], nspl=8, T1=[0, 0], T2=[0, 0], gate="cnot")
    elif name == "S888":  # three 8-level oscillators, Schroedinger, N = 512 (reference: sparse-matrix operator)
        p = dict(nlev=[8, 8, 8], lind="none", nt=ntime or 100, dt=0.02,
                 trans=[4.1, 4.3, 4.5], rot=[4.08, 4.27, 4.5], selfk=[0.01, 0.012, 0.008], crossk=[0.0005, 0.0002, 0.0004], J=[0.002, 0.0, 0.001],
                 car=[[0.0, -0.05], [0.0], [0.0, 0.04]], nspl=6, T1=[0, 0, 0], T2=[0, 0, 0], gate=None)
    elif name == "Q5L":  # five qubits, Lindblad, dim 1024
        p = dict(nlev=[2] * 5, lind="both", nt=ntime or 100, dt=0.05,
                 trans=[4.1, 4.2, 4.3, 4.4, 4.5], rot=[4.1, 4.2, 4.3, 4.4, 4.45], selfk=[0.2] * 5,
                 crossk=[0.001 * (i + 1) for i in range(10)], J=[0.0] * 10,
                 car=[[0.0]] * 5, nspl=5, T1=[50.0, 60.0, 0.0, 70.0, 80.0], T2=[30.0, 0.0, 35.0, 40.0, 0.0], gate=None)
    elif name == "G333":  # three qutrits, Lindblad, dim 729 (multi-element-per-thread kernels)
        p = dict(nlev=[3, 3, 3], lind="both", nt=ntime or 100, dt=0.05,
                 trans=[4.1, 4.3, 4.5], rot=[4.1, 4.25, 4.45], selfk=[0.2, 0.21, 0.22], crossk=[0.002, 0.001, 0.003],
                 J=[0.004, 0.0, 0.002], car=[[0.0, -0.2], [0.0], [0.0, -0.22]], nspl=7, T1=[60.0, 70.0, 0.0], T2=[30.0, 0.0, 35.0], gate=None)
    else:
        raise KeyError(name)
    p["name"] = name
    p["T"] = p["dt"] * p["nt"]
    return p


# level tuples the reference's matrix-free operator is instantiated for (src/mastereq.cpp:2977-3239); any other tuple
# runs through its sparse-matrix operator (same mathematics, usematfree = false)
REF_MATFREE_TUPLES = {(n,) for n in range(2, 11)} | {(3, 20), (3, 10), (1, 1), (2, 2), (3, 3), (4, 4), (5, 5), (10, 10), (20, 20),
                                                     (2, 2, 2), (2, 3, 4), (3, 3, 3), (4, 4, 4), (2, 2, 2, 2), (3, 3, 3, 3), (4, 4, 4, 4),
                                                     (2, 2, 2, 2, 2), (3, 3, 3, 3, 3)}


def ref_has_matfree(name):
    return tuple(system_params(name)["nlev"]) in REF_MATFREE_TUPLES


def cfg_system(name, ntime=None, linsolve="gmres", maxiter=20, flags=0, stepper="IMR"):
    """Returns (desc, ninit_full) for the named configuration (full `basis` set of initial conditions)."""
    p = system_params(name, ntime)
    nlev, Q = p["nlev"], len(p["nlev"])
    rot, trans = p["rot"], p["trans"]
    eta = []
    for i in range(Q):
        for j in range(i + 1, Q):
            eta.append(TWO_PI * (rot[i] - rot[j]))
    T = p["T"]
    desc = make_desc(
        nlevels=nlev, lindblad=p["lind"], detuning=[TWO_PI * (trans[k] - rot[k]) for k in range(Q)],
        selfkerr=[TWO_PI * s for s in p["selfk"]], crosskerr=[TWO_PI * c for c in p["crossk"]], Jkl=[TWO_PI * j for j in p["J"]], eta=eta,
        decay_time=p["T1"], dephase_time=p["T2"], ntime=p["nt"], total_time=T, linsolve=linsolve, maxiter=maxiter, stepper=stepper,
        segments=[[{"type": "spline", "nsplines": p["nspl"], "tstart": 0.0, "tstop": T}] for _ in range(Q)],
        carriers=[[TWO_PI * c for c in p["car"][k]] for k in range(Q)], flags=flags)
    N = int(np.prod(nlev))
    ninit = N * N if p["lind"] != "none" else N
    return desc, ninit


def cfg_text(name, ntime=None, linsolve="gmres", maxiter=20, runtype="gradient", params_file="params.dat", datadir="./data_out",
             objective="Jtrace", initialcondition="basis", usematfree=True, target=None, control_init=None, outputs="none", output_frequency=None,
             optim_maxiter=0, weights="1.0", penalty=0.0, penalty_param=0.0, penalty_energy=0.0, nessential=None, enforceBC=False, pipulse=None, timestepper="IMR", segments=None,
             penalty_variation=0.0, penalty_dpdm=0.0, overrides=None):
    """The same system as a reference .cfg (keys: config_template.cfg; units GHz/ns).  Outputs are switched off
     and the design vector is read from `params_file`."""
    p = system_params(name, ntime)
    Q = len(p["nlev"])
    j = lambda v: ",".join(repr(float(x)) if not isinstance(x, int) else str(x) for x in v)
    lines = [
        "rand_seed = 1234",
        "nlevels = " + ",".join(str(n) for n in p["nlev"]),
        "nessential = " + ",".join(str(n) for n in (nessential or p["nlev"])),
        f"ntime = {p['nt']}",
        f"dt = {p['dt']!r}",
        f"runtype = {runtype}",
        f"initialcondition = {initialcondition}",
        "transfreq = " + j(p["trans"]),
        "rotfreq = " + j(p["rot"]),
        "selfkerr = " + j(p["selfk"]),
        "crosskerr = " + (j(p["crossk"]) if p["crossk"] else "0.0"),
        "Jkl = " + (j(p["J"]) if p["J"] else "0.0"),
        f"collapse_type = {p['lind']}",
        "decay_time = " + j([float(x) for x in p["T1"]]),
        "dephase_time = " + j([float(x) for x in p["T2"]]),
        "control_enforceBC = " + ("true" if enforceBC else "false"),
        "usematfree = " + ("true" if usematfree else "false"),
        f"linearsolver_type = {linsolve}",
        f"linearsolver_maxiter = {maxiter}",
        f"timestepper = {timestepper}",
        f"datadir = {datadir}",
        f"output_frequency = {output_frequency or p['nt']}",
        "optim_monitor_frequency = 1",
        "optim_objective = " + objective,
        "optim_weights = " + weights,
        "optim_regul = 1e-05",
        f"optim_penalty = {penalty!r}",
        f"optim_penalty_param = {penalty_param!r}",
        f"optim_penalty_dpdm = {penalty_dpdm!r}",
        f"optim_penalty_energy = {penalty_energy!r}",
        f"optim_penalty_variation = {penalty_variation!r}",
        "optim_regul_tik0 = false",
        f"optim_maxiter = {optim_maxiter}",
    ]
    if pipulse:
        lines.append("apply_pipulse = " + pipulse)
    if target:
        lines.append("optim_target = " + target)
        lines.append("gate_rot_freq = " + j([0.0] * Q))
    elif p["gate"]:
        lines.append(f"optim_target = gate, {p['gate']}")
        lines.append("gate_rot_freq = " + j([0.0] * Q))
    else:
        lines.append("optim_target = pure, " + ",".join("0" for _ in range(Q)))
    for k in range(Q):
        lines.append(f"carrier_frequency{k} = " + j(p["car"][k]))
        lines.append(f"control_segments{k} = " + (segments or f"spline, {p['nspl']}"))
        lines.append(f"control_initialization{k} = " + (control_init or f"file, {params_file}"))
        lines.append(f"control_bounds{k} = 1e+12")
        lines.append(f"output{k} = {outputs}")
    for key, val in (overrides or {}).items():      # replace (or add) single settings, e.g. {"collapse_type": "decay"}
        lines = [ln for ln in lines if ln.split("=")[0].strip() != key] + [f"{key} = {val}"]
    return "\n".join(lines) + "\n"


def random_alpha(ndesign, seed=1234, amp=0.01):
    rng = np.random.default_rng(seed)
    return TWO_PI * amp * (2.0 * rng.random(ndesign) - 1.0)


def basis_states(N, lindblad):
    """The reference's `basis` initial conditions (src/optimtarget.cpp:605-690) in the [re | im] column-major layout."""
    if not lindblad:
        out = np.zeros((N, 2 * N))
        for i in range(N):
            out[i, i] = 1.0
        return out
    dim = N * N
    out = np.zeros((dim, 2 * dim))
    for iinit in range(dim):
        k, j = iinit % N, iinit // N
        if k == j:
            out[iinit, k + k * N] = 1.0
        elif k < j:
            out[iinit, k + k * N] = 0.5; out[iinit, j + j * N] = 0.5
            out[iinit, k + j * N] = 0.5; out[iinit, j + k * N] = 0.5
        else:
            out[iinit, k + k * N] = 0.5; out[iinit, j + j * N] = 0.5
            out[iinit, dim + k + j * N] = -0.5; out[iinit, dim + j + k * N] = 0.5
    return out
