"""CPU check of the mathematics behind the one-reduction GMRES of the sweep kernels (quandary_b200/csrc/qb_mono.cuh): the GMRES
iterate and its residual norm from the Gram matrix of the Neumann-scaled monomial Krylov basis u_j = (alpha M)^j b through an
L D L^T recurrence -- restated here in numpy line by line -- against the textbook definition (least squares over the Krylov space,
what Arnoldi/KSPGMRES computes), on the reference's own operator (oracle port) for BASELINE systems C3 and C4.
No GPU involved: this pins the algorithm, tests/test_engine_gpu.py pins the kernels."""
import numpy as np
import pytest

import qb_oracle
from quandary_b200 import desc_ndesign
from quandary_b200.systems import basis_states, cfg_system, random_alpha


def mono_gmres(G, maxiter, abstol=1e-10, reltol=1e-20, pivot_min=1e-6):
    """qb_mono.cuh: ldl_row (row by row), the stopping rule, mono_coeffs.  G: Gram matrix of u_0..u_s.  Returns (m, rho_m, c)."""
    n = G.shape[0]
    L = np.eye(n)
    Dinv = np.zeros(n)
    zp = np.zeros(n)
    sigma = 0.0
    tol2 = max(reltol * reltol * G[0, 0], abstol * abstol)
    if not G[0, 0] > tol2:
        return 0, np.sqrt(G[0, 0]), np.zeros(0)
    m = -1
    for r in range(n):
        t = np.zeros(r)
        d = G[r, r]
        for i in range(r):
            ti = G[r, i] - sum(t[k] * L[i, k] for k in range(i))
            t[i] = ti
            L[r, i] = ti * Dinv[i]
            d -= ti * L[r, i]
        assert d > pivot_min * G[r, r], "ill-conditioned basis: the kernels hand such a solve to the Arnoldi solver"
        Dinv[r] = 1.0 / d
        zp[r] = 1.0 - sum(L[r, i] * zp[i] for i in range(r))
        sigma = Dinv[r] if r == 0 else sigma + zp[r] * zp[r] * Dinv[r]
        if r >= 1 and (sigma * tol2 >= 1.0 or r == maxiter):
            m = r
            break
    assert m > 0, "not converged within the basis"
    e = np.zeros(m + 1)
    for j in range(m, -1, -1):
        e[j] = zp[j] * Dinv[j] - sum(L[i, j] * e[i] for i in range(j + 1, m + 1))
    rho2 = 1.0 / sigma
    c = np.array([e[j + 1:].sum() * rho2 for j in range(m)])
    return m, np.sqrt(rho2), c


@pytest.mark.parametrize("name,transpose", [("C3", False), ("C3", True), ("C4", False), ("C4", True)])
def test_gram_recurrence_gives_the_gmres_iterate_and_residual(name, transpose):
    orc = qb_oracle.load()
    desc, ninit = cfg_system(name, ntime=50)
    alpha = random_alpha(desc_ndesign(desc), amp=0.02)
    N = int(np.prod([desc.nlevels[k] for k in range(desc.nosc)]))
    dt = desc.total_time / desc.ntime
    a = dt / 2.0
    t = 7.5 * dt
    M = lambda x: orc.rhs_apply(desc, alpha, t, x[None, :], transpose=transpose)[0]
    rng = np.random.default_rng(5)
    for trial in range(3):
        x = basis_states(N, True)[rng.integers(ninit)] + 0.05 * rng.standard_normal(2 * N * N)
        b = M(x)
        U = [b]
        for _ in range(6):
            U.append(a * M(U[-1]))
        U = np.array(U)
        m, rho, c = mono_gmres(U @ U.T, maxiter=20)
        y = c @ U[:m]
        # textbook GMRES iterate of the same length: least squares over span{b, Ab, .., A^{m-1} b}, A = I - a M
        A = lambda v: v - a * M(v)
        K = U[:m].T
        AK = np.array([A(U[j]) for j in range(m)]).T
        cref, res, _, _ = np.linalg.lstsq(AK, b, rcond=None)
        yref = K @ cref
        rref = np.linalg.norm(b - AK @ cref)
        # (the textbook residual b - A K c is itself a difference of O(1) numbers: 1e-16 absolute on a residual of 1e-11)
        assert abs(rho - rref) <= 1e-4 * rref, (m, rho, rref)
        assert np.linalg.norm(y - yref) <= 1e-11 * np.linalg.norm(yref)
        # the stopping rule picked the first length that meets the reference's tolerance (atol 1e-10), as KSPConvergedDefault does
        assert rho <= 1e-10
        if m > 1:
            c1ref = np.linalg.lstsq(AK[:, :m - 1], b, rcond=None)[0]
            assert np.linalg.norm(b - AK[:, :m - 1] @ c1ref) > 1e-10
        # by-product used by the adjoint step: a M y = sum_j c_j u_{j+1}
        assert np.linalg.norm(a * M(y) - c @ U[1:m + 1]) <= 1e-12 * np.linalg.norm(y)


def test_gram_recurrence_respects_the_iteration_cap_and_the_zero_right_hand_side():
    rng = np.random.default_rng(0)
    Mx = rng.standard_normal((12, 12)) * 0.05
    b = rng.standard_normal(12)
    U = [b]
    for _ in range(4):
        U.append(Mx @ U[-1])
    U = np.array(U)
    m, rho, c = mono_gmres(U @ U.T, maxiter=2, abstol=1e-300)
    assert m == 2 and c.size == 2 and rho > 0      # capped: the iterate of length maxiter is returned unconverged, as KSP does
    m0, rho0, c0 = mono_gmres(np.zeros((3, 3)), maxiter=5)
    assert m0 == 0 and c0.size == 0                # b = 0: converged before the first iteration, y = 0
