"""Synthetic systems used by the parity tests (BASELINE.json configs C1-C4 at reduced ntime, plus generic-level cases)."""
import math

import numpy as np

from quandary_b200 import make_desc

TWO_PI = 2.0 * math.pi


def cfg_system(name, ntime=None, linsolve="gmres", maxiter=20, flags=0):
    """Returns (desc, ninit_full) for the named configuration.  Frequencies follow SURVEY.md section 8d."""
    if name == "C1":  # tests/regression/cnot/cnot.cfg: (2,2) Schroedinger, 3 carriers x 150 splines... here 10 splines
        nlev = [2, 2]; lind = "none"; nt = ntime or 1000; T = 0.1 * nt
        trans = [4.10595, 4.81526]; rot = [4.10595, 4.81526]; selfk = [0.2198, 0.2252]; crossk = [0.01]; J = [0.0]
        car = [[0.0, -0.01, 0.02], [0.0, -0.015, 0.01]]; nspl = 10; T1 = [0, 0]; T2 = [0, 0]
    elif name == "C2":
        nlev = [2, 2]; lind = "both"; nt = ntime or 1000; T = 0.1 * nt
        trans = [4.10595, 4.81526]; rot = [4.10595, 4.81526]; selfk = [0.2198, 0.2252]; crossk = [0.01]; J = [0.0]
        car = [[0.0, -0.01, 0.02], [0.0, -0.015, 0.01]]; nspl = 10; T1 = [56.0, 60.0]; T2 = [28.0, 30.0]
    elif name == "C3":
        nlev = [2, 2, 2]; lind = "both"; nt = ntime or 2000; T = 0.05 * nt
        trans = [4.10595, 4.81526, 5.2]; rot = trans; selfk = [0.2198, 0.2252, 0.21]; crossk = [0.1, 0.01, 0.1]; J = [0.0] * 3
        car = [[0.0, 0.05], [0.0, -0.05], [0.0, 0.03]]; nspl = 50; T1 = [80.0, 90.0, 100.0]; T2 = [40.0, 45.0, 50.0]
    elif name == "C4":
        nlev = [2, 2, 2, 2]; lind = "both"; nt = ntime or 10000; T = 0.01 * nt
        trans = [4.1, 4.2, 4.3, 4.4]; rot = trans; selfk = [0.2] * 4; crossk = [0.001] * 6; J = [0.0] * 6
        car = [[0.0, 0.01, -0.01]] * 4; nspl = 15; T1 = [100.0, 110.0, 120.0, 130.0]; T2 = [50.0, 55.0, 60.0, 65.0]
    elif name == "G32J":  # generic levels with Jaynes-Cummings coupling, Lindblad
        nlev = [3, 2]; lind = "both"; nt = ntime or 200; T = 0.1 * nt
        trans = [4.1, 4.8]; rot = [4.0, 4.6]; selfk = [0.2, 0.22]; crossk = [0.01]; J = [0.005]
        car = [[0.0, 0.2], [0.0]]; nspl = 8; T1 = [50.0, 60.0]; T2 = [0.0, 30.0]
    elif name == "G44J":  # two qudits, dim 256, coupling, decay only
        nlev = [4, 4]; lind = "decay"; nt = ntime or 100; T = 0.1 * nt
        trans = [4.1, 4.3]; rot = [4.05, 4.2]; selfk = [0.2, 0.21]; crossk = [0.002]; J = [0.003]
        car = [[0.0, -0.2], [0.0, 0.21]]; nspl = 6; T1 = [70.0, 80.0]; T2 = [0.0, 0.0]
    elif name == "G232S":  # Schroedinger, mixed levels, coupling
        nlev = [2, 3, 2]; lind = "none"; nt = ntime or 200; T = 0.1 * nt
        trans = [4.1, 4.3, 4.6]; rot = [4.1, 4.25, 4.5]; selfk = [0.2, 0.21, 0.22]; crossk = [0.002, 0.001, 0.003]; J = [0.004, 0.0, 0.002]
        car = [[0.0], [0.0, 0.1], [0.0]]; nspl = 7; T1 = [0] * 3; T2 = [0] * 3
    elif name == "G5":  # one oscillator, 5 levels, dephasing only
        nlev = [5]; lind = "dephase"; nt = ntime or 100; T = 0.1 * nt
        trans = [4.1]; rot = [4.0]; selfk = [0.2]; crossk = []; J = []
        car = [[0.0, -0.2, -0.4]]; nspl = 9; T1 = [0.0]; T2 = [20.0]
    elif name == "Q5":  # five qubits, Schroedinger (N = 32)
        nlev = [2] * 5; lind = "none"; nt = ntime or 100; T = 0.1 * nt
        trans = [4.1, 4.2, 4.3, 4.4, 4.5]; rot = [4.1, 4.2, 4.3, 4.4, 4.45]; selfk = [0.2] * 5; crossk = [0.001 * (i + 1) for i in range(10)]
        J = [0.0, 0.002, 0.0, 0.0, 0.001, 0.0, 0.0, 0.0, 0.003, 0.0]
        car = [[0.0]] * 5; nspl = 5; T1 = [0] * 5; T2 = [0] * 5
    else:
        raise KeyError(name)
    Q = len(nlev)
    eta = []
    for i in range(Q):
        for j in range(i + 1, Q):
            eta.append(TWO_PI * (rot[i] - rot[j]))
    desc = make_desc(
        nlevels=nlev, lindblad=lind, detuning=[TWO_PI * (trans[k] - rot[k]) for k in range(Q)],
        selfkerr=[TWO_PI * s for s in selfk], crosskerr=[TWO_PI * c for c in crossk], Jkl=[TWO_PI * j for j in J], eta=eta,
        decay_time=T1, dephase_time=T2, ntime=nt, total_time=T, linsolve=linsolve, maxiter=maxiter,
        segments=[[{"type": "spline", "nsplines": nspl, "tstart": 0.0, "tstop": T}] for _ in range(Q)],
        carriers=[[TWO_PI * c for c in car[k]] for k in range(Q)], flags=flags)
    N = int(np.prod(nlev))
    ninit = N * N if lind != "none" else N
    return desc, ninit


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
