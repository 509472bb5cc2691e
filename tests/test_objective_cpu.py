"""The numpy restatement of OptimTarget that checks the device objective kernels (tests/test_engine_gpu.py::_objective_numpy) is itself
pinned to the reference: applied to the final states the UNMODIFIED reference binary wrote (rho_Re/rho_Im files of the fixture
tests/golden/host/C2_simulation_outputs, 10 significant digits), it must reproduce the terminal cost and average fidelity of the
reference's own optim_history.dat.  CPU only."""
import os

import numpy as np

import quandary_b200 as qb
from quandary_b200.systems import basis_states
from test_engine_gpu import _objective_bar_numpy, _objective_numpy
from test_host_gpu import GOLD, load


def _final_states(case, ninit, dim):
    out = np.zeros((ninit, 2 * dim))
    for i in range(ninit):
        re = load(os.path.join(GOLD, case, "base", f"rho_Re.iinit{i:04d}.dat"))
        im = load(os.path.join(GOLD, case, "base", f"rho_Im.iinit{i:04d}.dat"))
        out[i, :dim], out[i, dim:] = re[-1, 1:], im[-1, 1:]
    return out


def test_numpy_objective_reproduces_reference_cost_and_fidelity():
    N, dim, ninit = 4, 16, 16
    V = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=float)       # cnot, src/gate.cpp:300-312
    rho0 = basis_states(N, True)
    xT = _final_states("C2_simulation_outputs", ninit, dim)
    hist = load(os.path.join(GOLD, "C2_simulation_outputs", "base", "optim_history.dat"))[0]
    fid_ref, cost_ref = hist[4], hist[5]
    jsum = fsum = 0.0
    for i in range(ninit):
        r0 = (rho0[i, :dim] + 1j * rho0[i, dim:]).reshape(N, N, order="F")
        tau_m = V @ r0 @ V.T
        tau = np.concatenate([tau_m.real.ravel(order="F"), tau_m.imag.ravel(order="F")])
        purity = float(np.sum(rho0[i] ** 2))
        jre, jim, hre, him = _objective_numpy(qb.OBJ_JTRACE, True, N, xT[i], tau, -1, purity)
        jsum += jre / ninit
        fsum += hre / ninit
    assert abs((1.0 - jsum) - cost_ref) < 5e-9 and abs(fsum - fid_ref) < 5e-9, (1.0 - jsum, cost_ref, fsum, fid_ref)


def test_numpy_objective_bar_is_the_derivative_of_the_objective():
    rng = np.random.default_rng(3)
    N = 3
    for lind in (True, False):
        dim = N * N if lind else N
        for kind, pure in [(qb.OBJ_JFROBENIUS, -1), (qb.OBJ_JFROBENIUS, 1), (qb.OBJ_JTRACE, -1), (qb.OBJ_JTRACE, 2), (qb.OBJ_JMEASURE, 0)]:
            x = rng.standard_normal(2 * dim)
            tau = rng.standard_normal(2 * dim) if pure < 0 else None
            rb, ib = 0.7, (-0.4 if not lind else 0.0)
            g = _objective_bar_numpy(kind, lind, N, x, tau, pure, 0.8, rb, ib)
            fd = np.zeros_like(x)
            for k in range(2 * dim):
                e = np.zeros_like(x)
                e[k] = 1e-6
                jp = _objective_numpy(kind, lind, N, x + e, tau, pure, 0.8)
                jm = _objective_numpy(kind, lind, N, x - e, tau, pure, 0.8)
                fd[k] = (rb * (jp[0] - jm[0]) + ib * (jp[1] - jm[1])) / 2e-6
            assert np.max(np.abs(g - fd)) < 1e-8, (lind, kind, pure, np.max(np.abs(g - fd)))
