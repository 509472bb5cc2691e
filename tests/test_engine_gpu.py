"""GPU parity tests: the CUDA engine (through the C ABI) against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

import quandary_b200 as qb
from quandary_b200 import Engine, desc_dim, desc_ndesign
from qb_systems import TWO_PI, basis_states, cfg_system, random_alpha

pytestmark = pytest.mark.gpu

SMALL = ["C1", "C2", "C3", "G32J", "G44J", "G232S", "G5", "Q5"]


def relerr(a, b):
    return np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b)))


@pytest.mark.parametrize("name", SMALL + ["C4"])
def test_rhs_apply_matches_oracle(oracle, name):
    desc, _ = cfg_system(name, ntime=10)
    alpha = random_alpha(desc_ndesign(desc), amp=0.05)
    eng = Engine(desc)
    eng.set_params(alpha)
    rng = np.random.default_rng(1234)
    x = rng.standard_normal((3, 2 * desc_dim(desc)))
    for t in (0.0, 0.37 * desc.total_time, desc.total_time):
        for tr in (False, True):
            y = eng.rhs_apply(t, x, transpose=tr)
            yo = oracle.rhs_apply(desc, alpha, t, x, transpose=tr)
            assert relerr(y, yo) < 1e-13, (name, t, tr)
    eng.close()


@pytest.mark.parametrize("name", ["C2", "G32J", "G5"])
def test_eval_controls_matches_oracle(oracle, name):
    desc, _ = cfg_system(name, ntime=10)
    alpha = random_alpha(desc_ndesign(desc), amp=0.05)
    eng = Engine(desc)
    eng.set_params(alpha)
    t = np.linspace(0.0, desc.total_time, 57)
    p, q = eng.eval_controls(t)
    for i, ti in enumerate(t):
        for k in range(desc.nosc):
            po, qo = oracle.eval_control(desc, alpha, k, float(ti))
            assert abs(p[i, k] - po) < 1e-13 * max(1.0, abs(po)) and abs(q[i, k] - qo) < 1e-13 * max(1.0, abs(qo))
    eng.close()


@pytest.mark.parametrize("linsolve", ["gmres", "neumann"])
@pytest.mark.parametrize("name", SMALL)
def test_forward_matches_oracle(oracle, name, linsolve):
    desc, ninit = cfg_system(name, ntime=60, linsolve=linsolve)
    alpha = random_alpha(desc_ndesign(desc), amp=0.02)
    N = int(np.prod([desc.nlevels[k] for k in range(desc.nosc)]))
    rho0 = basis_states(N, desc.lindblad != 0)
    sel = np.unique(np.linspace(0, ninit - 1, 6).astype(int))
    rho0 = rho0[sel]
    eng = Engine(desc)
    eng.set_params(alpha)
    rhoT = eng.forward(rho0, store=True)
    st = eng.stats()
    iters = evals = 0
    for i in range(rho0.shape[0]):
        rTo, traj, it, ev = oracle.solve_ode(desc, alpha, rho0[i], want_traj=True)
        iters += it; evals += ev
        assert relerr(rhoT[i], rTo) < 1e-11, (name, linsolve, i)
        for n in (0, 1, desc.ntime // 2, desc.ntime):
            assert relerr(eng.get_state(i, n), traj[n]) < 1e-11
    # same algorithm, same stopping rule => same work
    assert st["linsolve_iters"] == iters and st["rhs_evals"] == evals, (st, iters, evals)
    eng.close()


@pytest.mark.parametrize("flags", [0, qb.FLAG_RECOMPUTE_STAGE])
@pytest.mark.parametrize("linsolve", ["gmres", "neumann"])
@pytest.mark.parametrize("name", SMALL)
def test_adjoint_gradient_matches_oracle(oracle, name, linsolve, flags):
    desc, ninit = cfg_system(name, ntime=40, linsolve=linsolve, flags=flags)
    nd = desc_ndesign(desc)
    alpha = random_alpha(nd, amp=0.02)
    N = int(np.prod([desc.nlevels[k] for k in range(desc.nosc)]))
    rho0 = basis_states(N, desc.lindblad != 0)
    sel = np.unique(np.linspace(0, ninit - 1, 5).astype(int))
    rho0 = rho0[sel]
    rng = np.random.default_rng(7)
    rbar = rng.standard_normal(rho0.shape)
    eng = Engine(desc)
    eng.set_params(alpha)
    rhoT = eng.forward(rho0, store=True)
    grad, xb0 = eng.adjoint(rbar, want_rho0_bar=True)
    gsum = np.zeros(nd)
    for i in range(rho0.shape[0]):
        rTo, traj, _, _ = oracle.solve_ode(desc, alpha, rho0[i], want_traj=desc.lindblad != 0)
        go, xbo, _, _ = oracle.solve_adjoint(desc, alpha, rbar[i], rTo, traj)
        gsum += go
        assert relerr(xb0[i], xbo) < 1e-10, (name, i)
    # flags=0 replaces the re-solved stage by (x_{n-1}+x_n)/2: identical up to the linear-solver tolerance (1e-10)
    tol = 1e-10 if flags else 2e-9
    assert relerr(grad, gsum) < tol, (name, linsolve, flags, relerr(grad, gsum))
    eng.close()


def test_full_basis_c3_gradient(oracle):
    """BASELINE config 3 at reduced ntime: all 64 initial conditions, gradient summed in IC order."""
    desc, ninit = cfg_system("C3", ntime=50, flags=qb.FLAG_RECOMPUTE_STAGE)
    nd = desc_ndesign(desc)
    alpha = random_alpha(nd, amp=0.02)
    rho0 = basis_states(8, True)
    rng = np.random.default_rng(3)
    rbar = rng.standard_normal(rho0.shape) / ninit
    eng = Engine(desc)
    eng.set_params(alpha)
    rhoT = eng.forward(rho0, store=True)
    grad = eng.adjoint(rbar)
    rTo, go, _, _ = oracle.batch_gradient(desc, alpha, rho0, rbar)
    assert relerr(rhoT, rTo) < 1e-11
    assert relerr(grad, go) < 1e-10
    eng.close()


def test_c4_full_size_properties():
    """BASELINE config 4 at full width (256 initial conditions, N=16), size-independent properties:
    trace preservation, Hermiticity, linearity in rho0, and the adjoint identity <xbar_T, Phi x_0> = <Phi^T xbar_T, x_0>."""
    desc, ninit = cfg_system("C4", ntime=200)
    assert ninit == 256
    dim = desc_dim(desc)
    alpha = random_alpha(desc_ndesign(desc), amp=0.02)
    rho0 = basis_states(16, True)
    eng = Engine(desc)
    eng.set_params(alpha)
    rhoT = eng.forward(rho0, store=True)
    diag = np.arange(16) * 17
    assert np.max(np.abs(rhoT[:, diag].sum(axis=1) - 1.0)) < 1e-10          # Tr rho(T) = 1
    assert np.max(np.abs(rhoT[:, dim + diag])) < 1e-12                        # real diagonal
    re = rhoT[:, :dim].reshape(-1, 16, 16); im = rhoT[:, dim:].reshape(-1, 16, 16)
    assert np.max(np.abs(re - re.transpose(0, 2, 1))) < 1e-10 and np.max(np.abs(im + im.transpose(0, 2, 1))) < 1e-10
    rng = np.random.default_rng(5)
    rbar = rng.standard_normal(rho0.shape)
    grad, xb0 = eng.adjoint(rbar, want_rho0_bar=True)
    lhs = np.einsum("ij,ij->i", rbar, rhoT)
    rhs = np.einsum("ij,ij->i", xb0, rho0)
    assert np.max(np.abs(lhs - rhs)) < 1e-10 * max(1.0, np.max(np.abs(lhs)))
    # linearity: propagate a random combination of the basis
    w = rng.standard_normal(ninit)
    mix = (w[:, None] * rho0).sum(axis=0, keepdims=True)
    out = eng.forward(mix)
    assert np.max(np.abs(out[0] - (w[:, None] * rhoT).sum(axis=0))) < 1e-9
    assert np.all(np.isfinite(grad))
    eng.close()


def test_gradient_finite_difference():
    """d/dalpha of J = <rbar, rho(T)> by central differences vs the adjoint gradient (C2, all 16 ICs)."""
    desc, ninit = cfg_system("C2", ntime=100)
    nd = desc_ndesign(desc)
    alpha = random_alpha(nd, amp=0.02)
    rho0 = basis_states(4, True)
    rng = np.random.default_rng(11)
    rbar = rng.standard_normal(rho0.shape)
    eng = Engine(desc)
    eng.set_params(alpha)
    eng.forward(rho0, store=True)
    grad = eng.adjoint(rbar)
    eps = 1e-6
    for j in rng.choice(nd, size=6, replace=False):
        a = alpha.copy(); a[j] += eps; eng.set_params(a); jp = np.sum(rbar * eng.forward(rho0))
        a[j] -= 2 * eps; eng.set_params(a); jm = np.sum(rbar * eng.forward(rho0))
        fd = (jp - jm) / (2 * eps)
        assert abs(fd - grad[j]) < 1e-6 * max(1.0, abs(grad[j])), (j, fd, grad[j])
    eng.close()


def test_errors_are_loud():
    desc, _ = cfg_system("C2", ntime=10)
    eng = Engine(desc)
    with pytest.raises(qb.QBError):
        eng.forward(basis_states(4, True))          # params not set
    eng.set_params(random_alpha(desc_ndesign(desc)))
    eng.forward(basis_states(4, True), store=False)
    with pytest.raises(qb.QBError):
        eng.adjoint(basis_states(4, True))          # Lindblad adjoint without stored trajectory
    eng.close()


@pytest.mark.parametrize("name", ["C2", "C3", "G32J"])
def test_stage_from_trajectory_equals_recomputed_stage(name):
    """Lindblad: (x_{n-1}+x_n)/2 from the stored trajectory IS the implicit-midpoint stage x_{n-1}+dt/2*k up to rounding
    (x_n = x_{n-1}+dt*k was formed from the same k), so both adjoint modes give the same gradient to ~1e-13."""
    grads = []
    for flags in (0, qb.FLAG_RECOMPUTE_STAGE):
        desc, ninit = cfg_system(name, ntime=80, flags=flags)
        alpha = random_alpha(desc_ndesign(desc), amp=0.02)
        N = int(np.prod([desc.nlevels[k] for k in range(desc.nosc)]))
        rho0 = basis_states(N, True)
        rbar = np.random.default_rng(2).standard_normal(rho0.shape)
        eng = Engine(desc)
        eng.set_params(alpha)
        eng.forward(rho0, store=True)
        grads.append(eng.adjoint(rbar))
        eng.close()
    assert relerr(grads[0], grads[1]) < 1e-12, relerr(grads[0], grads[1])


def test_device_resident_evaluation_repeats_bitwise():
    """upload once, evaluate twice (the bench's device-resident loop): identical bits, and equal to the host-buffer path."""
    desc, ninit = cfg_system("C3", ntime=30)
    alpha = random_alpha(desc_ndesign(desc), amp=0.02)
    rho0 = basis_states(8, True)
    rbar = np.random.default_rng(9).standard_normal(rho0.shape)
    eng = Engine(desc)
    eng.set_params(alpha)
    eng.upload_states(rho0, 0)
    eng.upload_states(rbar, 1)
    out = []
    for _ in range(2):
        eng.forward_dev(ninit, store=True)
        eng.adjoint_dev(ninit)
        out.append((eng.download_states(ninit, 0), eng.download_grad()))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    rT = eng.forward(rho0, store=True)
    g = eng.adjoint(rbar)
    assert np.array_equal(rT, out[0][0]) and np.array_equal(g, out[0][1])
    st = eng.stats()
    assert st["n_fwd"] == 3 and st["n_adj"] == 3 and st["total_adj_ms"] > 0.0
    assert qb.measure_fp64_peak(0) > 1.0
    eng.close()


# ---- qubit-system kernels (qb_qubit.cuh): C3 (E=2), C4 (E=4 and E=8) -------------------------------------------------
def _fwd_adj_vs_oracle(oracle, desc, ninit, N, nsel, tol_state=1e-11, tol_grad=1e-10):
    nd = desc_ndesign(desc)
    alpha = random_alpha(nd, amp=0.02)
    rho0 = basis_states(N, True)[np.unique(np.linspace(0, ninit - 1, nsel).astype(int))]
    rbar = np.random.default_rng(7).standard_normal(rho0.shape)
    eng = Engine(desc)
    eng.set_params(alpha)
    rhoT = eng.forward(rho0, store=True)
    st_f = eng.stats()
    mid = eng.get_state(1, desc.ntime // 2)
    grad, xb0 = eng.adjoint(rbar, want_rho0_bar=True)
    eng.close()
    gsum = np.zeros(nd)
    iters = evals = 0
    for i in range(rho0.shape[0]):
        rTo, traj, it, ev = oracle.solve_ode(desc, alpha, rho0[i], want_traj=True)
        iters += it; evals += ev
        assert relerr(rhoT[i], rTo) < tol_state, i
        if i == 1:
            assert relerr(mid, traj[desc.ntime // 2]) < tol_state
        go, xbo, _, _ = oracle.solve_adjoint(desc, alpha, rbar[i], rTo, traj)
        gsum += go
        assert relerr(xb0[i], xbo) < tol_grad, i
    assert relerr(grad, gsum) < tol_grad, relerr(grad, gsum)
    return st_f, iters, evals


@pytest.mark.parametrize("linsolve", ["gmres", "neumann"])
@pytest.mark.parametrize("name,epe", [("C3", "0"), ("C4", "4"), ("C4", "2"), ("C4", "1")])
def test_qubit_kernels_match_oracle(oracle, monkeypatch, name, epe, linsolve):
    monkeypatch.setenv("QB_QUBIT_E", epe)
    desc, ninit = cfg_system(name, ntime=40, linsolve=linsolve, flags=qb.FLAG_RECOMPUTE_STAGE)
    N = int(np.prod([desc.nlevels[k] for k in range(desc.nosc)]))
    st, iters, evals = _fwd_adj_vs_oracle(oracle, desc, ninit, N, 5)
    assert st["linsolve_iters"] == iters and st["rhs_evals"] == evals, (st, iters, evals)


@pytest.mark.parametrize("name,epe", [("C3", "0"), ("C4", "4"), ("C4", "2"), ("C4", "1")])
def test_qubit_kernels_long_krylov_and_restart(oracle, monkeypatch, name, epe):
    """abstol far below reach: every solve runs to linearsolver_maxiter = 34 iterations: the one-reduction solver hands the
    solve back after its longest basis, the Arnoldi fallback (basis in global memory) runs it through one GMRES restart at 30
    (PETSc default)."""
    monkeypatch.setenv("QB_QUBIT_E", epe)
    desc, ninit = cfg_system(name, ntime=6, linsolve="gmres", maxiter=34, flags=qb.FLAG_RECOMPUTE_STAGE)
    desc.linsolve_abstol = 1e-300
    N = int(np.prod([desc.nlevels[k] for k in range(desc.nosc)]))
    st, iters, evals = _fwd_adj_vs_oracle(oracle, desc, ninit, N, 3, tol_state=1e-10, tol_grad=1e-9)
    # (a few solves stop before maxiter when the recurrence residual underflows rtol*|r0|; where exactly is a rounding matter)
    assert abs(st["linsolve_iters"] - iters) <= 0.02 * iters and iters > 0.9 * 3 * 6 * 34


@pytest.mark.parametrize("name", ["C3", "C4"])
def test_qubit_kernels_agree_with_generic_kernels(monkeypatch, name):
    """the specialised and the general kernels are two implementations of the same arithmetic"""
    desc, ninit = cfg_system(name, ntime=50)
    N = int(np.prod([desc.nlevels[k] for k in range(desc.nosc)]))
    alpha = random_alpha(desc_ndesign(desc), amp=0.02)
    rho0 = basis_states(N, True)[:: max(1, ninit // 16)]
    rbar = np.random.default_rng(1).standard_normal(rho0.shape)
    res = []
    for mode in ("generic", "qubit"):
        monkeypatch.setenv("QB_KERNELS", mode)
        eng = Engine(desc)
        eng.set_params(alpha)
        rT = eng.forward(rho0, store=True)
        g = eng.adjoint(rbar)
        res.append((rT, g))
        eng.close()
    assert relerr(res[1][0], res[0][0]) < 1e-12 and relerr(res[1][1], res[0][1]) < 1e-11


@pytest.mark.parametrize("linsolve", ["gmres", "neumann"])
@pytest.mark.parametrize("name,epe", [("C3", "0"), ("C4", "4")])
def test_stored_stage_is_the_recomputed_stage(monkeypatch, name, epe, linsolve):
    """QB_FLAG_STORE_STAGE: the forward sweep keeps x_{n-1}+dt/2*k; the adjoint gradient is the one of the re-solving mode
    (ImplMidpoint::evolveBWD, src/timestepper.cpp:674-680) to rounding (the re-solve may build its Krylov basis in blocks of a
    different length, which changes the association of the reduction sums), and get_state still returns x_n."""
    monkeypatch.setenv("QB_QUBIT_E", epe)
    out = []
    for flags in (qb.FLAG_RECOMPUTE_STAGE, qb.FLAG_STORE_STAGE):
        desc, ninit = cfg_system(name, ntime=37, linsolve=linsolve, flags=flags)
        N = int(np.prod([desc.nlevels[k] for k in range(desc.nosc)]))
        alpha = random_alpha(desc_ndesign(desc), amp=0.02)
        rho0 = basis_states(N, True)[:: max(1, ninit // 8)]
        rbar = np.random.default_rng(4).standard_normal(rho0.shape)
        eng = Engine(desc)
        eng.set_params(alpha)
        rT = eng.forward(rho0, store=True)
        mid = eng.get_state(2, 20)
        g, xb = eng.adjoint(rbar, want_rho0_bar=True)
        out.append((rT, mid, g, xb))
        eng.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    for a, b in zip(out[0][2:], out[1][2:]):
        assert relerr(a, b) < 1e-13, relerr(a, b)


@pytest.mark.parametrize("name,epe", [("C3", "0"), ("C4", "4"), ("C4", "2"), ("C4", "1")])
def test_one_reduction_solver_equals_arnoldi(monkeypatch, name, epe):
    """The qubit-system kernels compute every GMRES iterate from the Gram matrix of the monomial Krylov basis (one reduction per
    solve); QB_FLAG_ARNOLDI runs the same solves with the step-by-step Arnoldi process (what KSPGMRES does).  Same iterates: equal
    iteration counts, states and gradients to rounding -- over enough steps that sparse initial states and dense late states,
    converged-at-2 and converged-at-4 solves all occur."""
    monkeypatch.setenv("QB_QUBIT_E", epe)
    out = []
    for flags in (qb.FLAG_STORE_STAGE, qb.FLAG_STORE_STAGE | qb.FLAG_ARNOLDI):
        desc, ninit = cfg_system(name, ntime=300, flags=flags)
        N = int(np.prod([desc.nlevels[k] for k in range(desc.nosc)]))
        alpha = random_alpha(desc_ndesign(desc), amp=0.02)
        rho0 = basis_states(N, True)[:: max(1, ninit // 16)]
        rbar = np.random.default_rng(4).standard_normal(rho0.shape)
        eng = Engine(desc)
        eng.set_params(alpha)
        rT = eng.forward(rho0, store=True)
        g, xb = eng.adjoint(rbar, want_rho0_bar=True)
        st = eng.stats()
        out.append((rT, g, xb, st["linsolve_iters"], st["rhs_evals"]))
        eng.close()
    assert relerr(out[0][0], out[1][0]) < 1e-12 and relerr(out[0][1], out[1][1]) < 1e-11 and relerr(out[0][2], out[1][2]) < 1e-12
    assert abs(out[0][3] - out[1][3]) <= 1e-3 * out[1][3] and abs(out[0][4] - out[1][4]) <= 1e-3 * out[1][4], (out[0][3:], out[1][3:])


@pytest.mark.parametrize("name", ["C3", "C4"])
def test_stages_only_storage(name):
    """store_trajectory = 2: the forward sweep keeps the stages only (half the trajectory memory); the adjoint sweep gives the
    same gradient bit for bit, the state trajectory is not available."""
    desc, ninit = cfg_system(name, ntime=37, flags=qb.FLAG_STORE_STAGE)
    N = int(np.prod([desc.nlevels[k] for k in range(desc.nosc)]))
    alpha = random_alpha(desc_ndesign(desc), amp=0.02)
    rho0 = basis_states(N, True)[:: max(1, ninit // 8)]
    rbar = np.random.default_rng(4).standard_normal(rho0.shape)
    eng = Engine(desc)
    eng.set_params(alpha)
    out = []
    for store in (qb.STORE_STATES, qb.STORE_ADJOINT_INPUT):
        rT = eng.forward(rho0, store=store)
        if store == qb.STORE_ADJOINT_INPUT:
            with pytest.raises(qb.QBError):
                eng.get_state(0, 5)
        out.append((rT, eng.adjoint(rbar)))
    eng.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


# ---- large-state kernels (qb_big.cuh): dim 729 (E=1), 1024 (E=1), 3600 and 4096 (E=4) -------------------------------
BIG = [("G333", 8), ("Q5L", 8), ("A320", 6), ("C5", 5)]


@pytest.mark.parametrize("name,nt", BIG)
def test_big_rhs_apply_matches_oracle(oracle, name, nt):
    desc, _ = cfg_system(name, ntime=nt)
    alpha = random_alpha(desc_ndesign(desc), amp=0.05)
    eng = Engine(desc)
    eng.set_params(alpha)
    x = np.random.default_rng(1234).standard_normal((2, 2 * desc_dim(desc)))
    for t in (0.0, 0.37 * desc.total_time):
        for tr in (False, True):
            y = eng.rhs_apply(t, x, transpose=tr)
            yo = oracle.rhs_apply(desc, alpha, t, x, transpose=tr)
            assert relerr(y, yo) < 1e-13, (name, t, tr, relerr(y, yo))
    eng.close()


def test_tensor_core_operator_matches_stencil_and_oracle(oracle, monkeypatch):
    """N = 64 (C5): the dense formulation -i[H, rho] + dissipator on the fp64 tensor pipe (qb_dense.cu, mma.sync m8n8k4 = DMMA) against
    the oracle and the stencil kernel on random (non-Hermitian) states"""
    desc, _ = cfg_system("C5", ntime=8)
    alpha = random_alpha(desc_ndesign(desc), amp=0.05)
    x = np.random.default_rng(1234).standard_normal((3, 2 * desc_dim(desc)))
    eng = Engine(desc)
    eng.set_params(alpha)
    for t in (0.0, 0.37 * desc.total_time):
        yo = oracle.rhs_apply(desc, alpha, t, x, transpose=False)
        monkeypatch.setenv("QB_DENSE_APPLY", "0")
        ys = eng.rhs_apply(t, x)
        monkeypatch.setenv("QB_DENSE_APPLY", "1")
        yd = eng.rhs_apply(t, x)
        assert relerr(ys, yo) < 1e-13 and relerr(yd, yo) < 1e-13, (relerr(ys, yo), relerr(yd, yo))
        assert not np.array_equal(ys, yd)          # two different kernels did run
    assert qb.measure_fp64_mma_peak(0) > 1.0
    eng.close()


@pytest.mark.parametrize("name,nt", [("G333", 30), ("Q5L", 30), ("A320", 12), ("C5", 25)])
def test_big_one_reduction_solver_equals_arnoldi(name, nt):
    """large-state kernels: the one-reduction GMRES (Gram rows accumulated while the basis is built, scalar solve by one warp) against
    the step-by-step Arnoldi solver (QB_FLAG_ARNOLDI): same iterates, hence equal iteration counts, states and gradients to rounding"""
    out = []
    for flags in (qb.FLAG_STORE_STAGE, qb.FLAG_STORE_STAGE | qb.FLAG_ARNOLDI):
        desc, ninit = cfg_system(name, ntime=nt, flags=flags)
        N = int(np.prod([desc.nlevels[k] for k in range(desc.nosc)]))
        alpha = random_alpha(desc_ndesign(desc), amp=0.02)
        rho0 = basis_states(N, True)[:: max(1, ninit // 6)][:6]
        rbar = np.random.default_rng(4).standard_normal(rho0.shape)
        eng = Engine(desc)
        eng.set_params(alpha)
        rT = eng.forward(rho0, store=True)
        g, xb = eng.adjoint(rbar, want_rho0_bar=True)
        st = eng.stats()
        out.append((rT, g, xb, st["linsolve_iters"], st["rhs_evals"]))
        eng.close()
    assert relerr(out[0][0], out[1][0]) < 1e-12 and relerr(out[0][1], out[1][1]) < 1e-10 and relerr(out[0][2], out[1][2]) < 1e-11
    assert abs(out[0][3] - out[1][3]) <= 2e-2 * out[1][3], (out[0][3:], out[1][3:])


@pytest.mark.parametrize("flags", [qb.FLAG_STORE_STAGE, qb.FLAG_RECOMPUTE_STAGE, 0])
@pytest.mark.parametrize("linsolve", ["gmres", "neumann"])
@pytest.mark.parametrize("name,nt", BIG)
def test_big_forward_adjoint_match_oracle(oracle, name, nt, linsolve, flags):
    if flags == 0 and linsolve == "neumann":
        pytest.skip("covered by the gmres variant")
    desc, ninit = cfg_system(name, ntime=nt, linsolve=linsolve, flags=flags)
    N = int(np.prod([desc.nlevels[k] for k in range(desc.nosc)]))
    nd = desc_ndesign(desc)
    alpha = random_alpha(nd, amp=0.02)
    sel = np.unique(np.linspace(0, ninit - 1, 3).astype(int))
    rho0 = np.zeros((sel.size, 2 * N * N))
    for r, ii in enumerate(sel):      # the reference's basis elements B_kj without building all N^2 of them
        k, j = ii % N, ii // N
        if k == j:
            rho0[r, k + k * N] = 1.0
        else:
            rho0[r, k + k * N] = 0.5; rho0[r, j + j * N] = 0.5
            if k < j:
                rho0[r, k + j * N] = 0.5; rho0[r, j + k * N] = 0.5
            else:
                rho0[r, N * N + k + j * N] = -0.5; rho0[r, N * N + j + k * N] = 0.5
    rbar = np.random.default_rng(7).standard_normal(rho0.shape)
    eng = Engine(desc)
    eng.set_params(alpha)
    rhoT = eng.forward(rho0, store=True)
    st = eng.stats()
    mid = eng.get_state(1, nt // 2)
    grad, xb0 = eng.adjoint(rbar, want_rho0_bar=True)
    eng.close()
    gsum = np.zeros(nd)
    iters = 0
    for i in range(rho0.shape[0]):
        rTo, traj, it, _ = oracle.solve_ode(desc, alpha, rho0[i], want_traj=True)
        iters += it
        assert relerr(rhoT[i], rTo) < 1e-11, (name, i, relerr(rhoT[i], rTo))
        if i == 1:
            assert relerr(mid, traj[nt // 2]) < 1e-11
        go, xbo, _, _ = oracle.solve_adjoint(desc, alpha, rbar[i], rTo, traj)
        gsum += go
        assert relerr(xb0[i], xbo) < 1e-10, (name, i)
    assert st["linsolve_iters"] == iters
    assert relerr(grad, gsum) < (2e-9 if flags == 0 else 1e-10), relerr(grad, gsum)


# ---- compositional steppers IMR4 / IMR8 (CompositionalImplMidpoint, src/timestepper.cpp:729-826) ----------------------
@pytest.mark.parametrize("stepper", ["IMR4", "IMR8"])
@pytest.mark.parametrize("name,nt", [("C2", 30), ("C3", 20), ("G32J", 24), ("C1", 30), ("G333", 4)])
def test_compositional_steppers_match_oracle(oracle, name, nt, stepper):
    for flags in (qb.FLAG_STORE_STAGE, qb.FLAG_RECOMPUTE_STAGE):
        desc, ninit = cfg_system(name, ntime=nt, flags=flags, stepper=stepper)
        N = int(np.prod([desc.nlevels[k] for k in range(desc.nosc)]))
        lind = desc.lindblad != 0
        nd = desc_ndesign(desc)
        alpha = random_alpha(nd, amp=0.02)
        rho0 = basis_states(N, lind)[np.unique(np.linspace(0, ninit - 1, 3).astype(int))]
        rbar = np.random.default_rng(7).standard_normal(rho0.shape)
        eng = Engine(desc)
        eng.set_params(alpha)
        rhoT = eng.forward(rho0, store=True)
        mid = eng.get_state(1, nt // 2)
        grad = eng.adjoint(rbar)
        eng.close()
        gsum = np.zeros(nd)
        for i in range(rho0.shape[0]):
            rTo, traj, _, _ = oracle.solve_ode(desc, alpha, rho0[i], want_traj=True)
            assert relerr(rhoT[i], rTo) < 1e-11, (name, stepper, i, relerr(rhoT[i], rTo))
            if i == 1:
                assert relerr(mid, traj[nt // 2]) < 1e-11       # getState(n) = the state after n full time steps
            go, _, _, _ = oracle.solve_adjoint(desc, alpha, rbar[i], rTo, traj if lind else None)
            gsum += go
        assert relerr(grad, gsum) < 1e-10, (name, stepper, flags, relerr(grad, gsum))


@pytest.mark.parametrize("nt", [1, 2, 3, 5])
@pytest.mark.parametrize("name", ["C3", "C4", "C2", "G333"])
def test_short_sweeps(oracle, name, nt):
    """sweeps shorter than the TMA ring depths (control-record ring: 4 slots, trajectory ring: 2 slots)"""
    for flags in (qb.FLAG_STORE_STAGE, qb.FLAG_RECOMPUTE_STAGE, 0):
        desc, ninit = cfg_system(name, ntime=nt, flags=flags)
        N = int(np.prod([desc.nlevels[k] for k in range(desc.nosc)]))
        nd = desc_ndesign(desc)
        alpha = random_alpha(nd, amp=0.02)
        rho0 = basis_states(N, True)[[0, ninit // 2 + 1]]
        rbar = np.random.default_rng(3).standard_normal(rho0.shape)
        eng = Engine(desc)
        eng.set_params(alpha)
        rhoT = eng.forward(rho0, store=True)
        grad = eng.adjoint(rbar)
        eng.close()
        gsum = np.zeros(nd)
        for i in range(2):
            rTo, traj, _, _ = oracle.solve_ode(desc, alpha, rho0[i], want_traj=True)
            assert relerr(rhoT[i], rTo) < 1e-12
            go, _, _, _ = oracle.solve_adjoint(desc, alpha, rbar[i], rTo, traj)
            gsum += go
        assert relerr(grad, gsum) < (1e-9 if flags == 0 else 1e-11), (name, nt, flags)


# ---- randomised systems: ragged level tuples, partial dissipation, mixed control bases --------------------------------
def _random_system(rng):
    Q = int(rng.integers(1, 4))
    nlev = [int(rng.integers(2, 5)) for _ in range(Q)]
    lind = ["none", "decay", "dephase", "both"][int(rng.integers(0, 4))]
    N = int(np.prod(nlev))
    if lind != "none" and N * N > 256 and rng.random() < 0.5:
        nlev = nlev[:-1] if Q > 1 else [3]
        Q = len(nlev)
        N = int(np.prod(nlev))
    npairs = Q * (Q - 1) // 2
    T = float(rng.uniform(0.5, 3.0))
    ntime = int(rng.integers(3, 25))
    segments, carriers = [], []
    for k in range(Q):
        kind = rng.integers(0, 4)
        if kind == 0:
            segs = [{"type": "spline", "nsplines": int(rng.integers(4, 12)), "tstart": 0.0, "tstop": T, "enforce_bc": bool(rng.integers(0, 2))}]
        elif kind == 1:
            segs = [{"type": "spline0", "nsplines": int(rng.integers(2, 9)), "tstart": 0.0, "tstop": T, "enforce_bc": False}]
        elif kind == 2:   # two segments, the second one of the other kind
            tm = T * float(rng.uniform(0.3, 0.7))
            segs = [{"type": "spline", "nsplines": int(rng.integers(4, 8)), "tstart": 0.0, "tstop": tm},
                    {"type": "spline0", "nsplines": int(rng.integers(2, 6)), "tstart": tm, "tstop": T}]
        else:
            segs = []        # non-controllable oscillator
        segments.append(segs)
        carriers.append([TWO_PI * float(f) for f in rng.uniform(-0.3, 0.3, size=int(rng.integers(1, 4)))])
    flags = [0, qb.FLAG_RECOMPUTE_STAGE, qb.FLAG_STORE_STAGE][int(rng.integers(0, 3))]
    maxiter = int(rng.integers(8, 25))
    if flags != qb.FLAG_RECOMPUTE_STAGE:
        # the stored / averaged stage equals the reference's recomputed one only up to the linear-solve residual, so these two
        # modes get a step size and an iteration cap at which the solves converge; the recompute mode (the reference's exact
        # order of operations) also runs the harsh draws where GMRES / Neumann stop at the iteration cap
        ntime = max(ntime, int(np.ceil(T / 0.08)))
        maxiter = 30
    desc = qb.make_desc(
        nlevels=nlev, lindblad=lind, detuning=list(TWO_PI * rng.uniform(-0.1, 0.1, Q)), selfkerr=list(TWO_PI * rng.uniform(0.0, 0.3, Q)),
        crosskerr=list(TWO_PI * rng.uniform(0.0, 0.02, npairs)), Jkl=[TWO_PI * float(j) if rng.random() < 0.6 else 0.0 for j in rng.uniform(0.0, 0.01, npairs)],
        eta=list(TWO_PI * rng.uniform(-0.5, 0.5, npairs)), decay_time=[float(t) if rng.random() < 0.7 else 0.0 for t in rng.uniform(5.0, 50.0, Q)],
        dephase_time=[float(t) if rng.random() < 0.7 else 0.0 for t in rng.uniform(5.0, 50.0, Q)], ntime=ntime, total_time=T,
        linsolve=["gmres", "neumann"][int(rng.integers(0, 2))], maxiter=maxiter, segments=segments, carriers=carriers, flags=flags)
    return desc, N, lind != "none"


@pytest.mark.parametrize("seed", range(32))
def test_random_systems_match_oracle(oracle, seed):
    rng = np.random.default_rng(9000 + seed)
    desc, N, lind = _random_system(rng)
    nd = desc_ndesign(desc)
    dim = desc_dim(desc)
    alpha = TWO_PI * 0.02 * (2.0 * rng.random(max(nd, 1)) - 1.0)
    alpha = alpha[:nd]
    nic = 3
    rho0 = rng.standard_normal((nic, 2 * dim))          # arbitrary (not even Hermitian) states: the sweeps are linear maps
    rbar = rng.standard_normal((nic, 2 * dim))
    eng = Engine(desc)
    eng.set_params(alpha)
    rhoT = eng.forward(rho0, store=True)
    grad, xb0 = eng.adjoint(rbar, want_rho0_bar=True)
    eng.close()
    gsum = np.zeros(nd)
    for i in range(nic):
        rTo, traj, _, _ = oracle.solve_ode(desc, alpha, rho0[i], want_traj=lind)
        assert relerr(rhoT[i], rTo) < 1e-11, (seed, i, relerr(rhoT[i], rTo))
        go, xbo, _, _ = oracle.solve_adjoint(desc, alpha, rbar[i], rTo, traj)
        gsum += go
        assert relerr(xb0[i], xbo) < 1e-9, (seed, i, relerr(xb0[i], xbo))
    if nd > 0 and np.max(np.abs(gsum)) > 0:
        assert relerr(grad, gsum) < (5e-9 if desc.flags == 0 else 1e-10), (seed, desc.flags, relerr(grad, gsum))


def test_more_loud_errors():
    desc, _ = cfg_system("C3", ntime=8)
    eng = Engine(desc)
    with pytest.raises(qb.QBError):
        eng.set_params(np.zeros(desc_ndesign(desc) + 1))        # wrong design-vector length
    eng.set_params(random_alpha(desc_ndesign(desc)))
    rho0 = basis_states(8, True)[:4]
    eng.forward(rho0, store=True)
    with pytest.raises(qb.QBError):
        eng.adjoint(rho0[:3])                                   # batch size differs from the forward sweep
    with pytest.raises(qb.QBError):
        eng.get_state(7, 0)                                     # initial condition out of range
    with pytest.raises(qb.QBError):
        eng.get_state(0, 99)                                    # time index out of range
    with pytest.raises(qb.QBError):
        eng.penalty_integrals(4)                                # no penalty installed
    eng.close()
    bad, _ = cfg_system("C3", ntime=8)
    bad.nlevels[0] = 40                                         # dim 25600 > 4096
    with pytest.raises(qb.QBError):
        Engine(bad)
    bad2, _ = cfg_system("C3", ntime=8)
    bad2.stepper = 2                                            # not a stepper
    with pytest.raises(qb.QBError):
        Engine(bad2)
    bad3, _ = cfg_system("C5", ntime=8)
    bad3.stepper = 1                                            # explicit Euler exists for state dimensions <= 256 only
    with pytest.raises(qb.QBError):
        Engine(bad3)


def test_explicit_euler_adjoint_is_the_derivative_of_its_forward_sweep_up_to_first_order():
    """timestepper = EE (ExplEuler, src/timestepper.cpp:484-519) on C3 through the general kernels: the forward sweep is x <- x + dt A(t_n) x
    exactly (checked against a numpy loop over qb_engine_rhs_apply), and the reduced gradient agrees with finite differences of the forward
    sweep to the O(dt) the reference's adjoint (A(t_n) instead of A(t_{n-1}) in the backward step) implies."""
    desc, ninit = cfg_system("C3", ntime=30)
    desc.stepper = 1
    nd = desc_ndesign(desc)
    alpha = random_alpha(nd, amp=0.02)
    rho0 = basis_states(8, True)[[0, 9, 37]]
    eng = Engine(desc)
    eng.set_params(alpha)
    rT = eng.forward(rho0, store=True)
    dt = desc.total_time / desc.ntime
    x = rho0.copy()
    for n in range(desc.ntime):
        x = x + dt * eng.rhs_apply(n * dt, x)
    assert relerr(rT, x) < 1e-13
    rbar = np.random.default_rng(3).standard_normal(rho0.shape)
    g = eng.adjoint(rbar)
    fd = np.zeros(6)
    idx = np.random.default_rng(5).choice(nd, 6, replace=False)
    for j, i in enumerate(idx):
        h = 1e-6
        ap, am = alpha.copy(), alpha.copy()
        ap[i] += h; am[i] -= h
        eng.set_params(ap); fp = float(np.sum(eng.forward(rho0) * rbar))
        eng.set_params(am); fm = float(np.sum(eng.forward(rho0) * rbar))
        fd[j] = (fp - fm) / (2 * h)
    eng.close()
    assert np.max(np.abs(g[idx] - fd)) < 0.3 * np.max(np.abs(fd)) + 1e-8, (g[idx], fd)


# ---- final-time objective on the device against a numpy restatement of OptimTarget ---------------------
def _objective_numpy(kind, lind, N, x, tau, pure_id, purity):
    """evalJ (src/optimtarget.cpp:713-797) and HilbertSchmidtOverlap(.., false) (:343-402) of ONE state in the reference
    layout [re | im]; returns (J_re, J_im, HS_re, HS_im)."""
    dim = x.size // 2
    re, im = x[:dim], x[dim:]
    idm = (pure_id + pure_id * N if lind else pure_id) if pure_id >= 0 else None
    if pure_id >= 0:
        hs = (re[idm], im[idm])
    elif lind:
        hs = (float(tau @ x), 0.0)
    else:
        tr, ti = tau[:dim], tau[dim:]
        hs = (float(tr @ re + ti @ im), float(-ti @ re + tr @ im))
    if kind == qb.OBJ_JFROBENIUS:
        ref = tau if pure_id < 0 else np.eye(1, 2 * dim, idm).ravel()
        return 0.5 * float(np.sum((x - ref) ** 2)), 0.0, hs[0], hs[1]
    if kind == qb.OBJ_JTRACE:
        return hs[0] / purity, hs[1], hs[0], hs[1]
    lam = np.abs(np.arange(N) - pure_id).astype(float)
    pop = re[np.arange(N) * (N + 1)] if lind else re ** 2 + im ** 2
    return float(lam @ pop), 0.0, hs[0], hs[1]


def _objective_bar_numpy(kind, lind, N, x, tau, pure_id, purity, rb, ib):
    """evalJ_diff (:800-862) seeded with (rb, ib)."""
    dim = x.size // 2
    out = np.zeros_like(x)
    idm = (pure_id + pure_id * N if lind else pure_id) if pure_id >= 0 else None
    if kind == qb.OBJ_JFROBENIUS:
        if pure_id < 0:
            return rb * (x - tau)
        out = rb * x
        out[idm] -= rb
        return out
    if kind == qb.OBJ_JTRACE:
        if pure_id >= 0:
            out[idm] = rb / purity
            out[idm + dim] = ib
        elif lind:
            out = rb / purity * tau
        else:
            tr, ti = tau[:dim], tau[dim:]
            out[:dim] = tr * rb / purity - ti * ib
            out[dim:] = ti * rb / purity + tr * ib
        return out
    lam = np.abs(np.arange(N) - pure_id).astype(float)
    if lind:
        out[np.arange(N) * (N + 1)] = lam * rb
    else:
        out[:dim] = 2.0 * rb * lam * x[:dim]
        out[dim:] = 2.0 * rb * lam * x[dim:]
    return out


@pytest.mark.parametrize("name,kind,pure", [("C3", qb.OBJ_JTRACE, -1), ("C3", qb.OBJ_JFROBENIUS, -1), ("C3", qb.OBJ_JFROBENIUS, 5), ("C3", qb.OBJ_JTRACE, 3),
                                            ("C3", qb.OBJ_JMEASURE, 2), ("C1", qb.OBJ_JTRACE, -1), ("C1", qb.OBJ_JFROBENIUS, -1), ("C1", qb.OBJ_JMEASURE, 1),
                                            ("C1", qb.OBJ_JTRACE, 2), ("G33", qb.OBJ_JTRACE, -1), ("C5", qb.OBJ_JMEASURE, 0), ("C5", qb.OBJ_JTRACE, -1)])
def test_device_objective_matches_numpy_restatement(oracle, name, kind, pure):
    ntime = 4 if name == "C5" else 20
    desc, _ = cfg_system(name, ntime=ntime)
    lind = desc.lindblad != 0
    N = int(np.prod([desc.nlevels[k] for k in range(desc.nosc)]))
    dim = desc_dim(desc)
    nd = desc_ndesign(desc)
    rng = np.random.default_rng(77)
    nic = 7 if name != "C5" else 3
    rho0 = rng.standard_normal((nic, 2 * dim)) / np.sqrt(dim)
    tau = rng.standard_normal((nic, 2 * dim)) / np.sqrt(dim) if pure < 0 else None
    purity = rng.uniform(0.5, 1.0, nic)
    w = rng.uniform(0.5, 1.5, nic)
    w /= w.sum()
    eng = Engine(desc)
    eng.set_params(random_alpha(nd, amp=0.02))
    eng.upload_states(rho0, 0)
    eng.forward_dev(nic, store=True)
    xT = eng.download_states(nic, 0)
    eng.set_objective(kind, w, 1.0 / 16, pure_id=pure, target=tau, purity=purity)
    got = eng.objective(nic)
    want = np.zeros(4)
    for i in range(nic):
        j = _objective_numpy(kind, lind, N, xT[i], None if tau is None else tau[i], pure, purity[i])
        want += np.array([w[i] * j[0], w[i] * j[1], j[2] / 16, j[3] / 16])
    assert np.max(np.abs(got - want)) <= 1e-13 * max(1.0, np.max(np.abs(want))), (got, want)
    rb, ib = -1.3, (0.7 if not lind else 0.0)
    eng.objective_bar(nic, rb, ib)
    grad = None
    eng.adjoint_dev(nic)
    grad = eng.download_grad()
    # the same terminal adjoints through the host-buffer call
    rbar = np.stack([_objective_bar_numpy(kind, lind, N, xT[i], None if tau is None else tau[i], pure, purity[i], w[i] * rb, w[i] * ib) for i in range(nic)])
    eng.forward_dev(nic, store=True)
    grad_host = eng.adjoint(rbar)
    eng.close()
    assert np.max(np.abs(grad - grad_host)) <= 1e-12 * max(1e-300, np.max(np.abs(grad_host))), np.max(np.abs(grad - grad_host))


def test_device_objective_loud_errors():
    desc, _ = cfg_system("C2", ntime=5)
    eng = Engine(desc)
    eng.set_params(random_alpha(desc_ndesign(desc)))
    with pytest.raises(qb.QBError):
        eng.objective(4)                                                   # nothing installed
    with pytest.raises(qb.QBError):
        eng.set_objective(qb.OBJ_JMEASURE, np.ones(4) / 4, 0.25, pure_id=-1, target=np.zeros((4, 32)))   # Jmeasure needs a pure target
    with pytest.raises(qb.QBError):
        eng.set_objective(qb.OBJ_JTRACE, np.ones(4) / 4, 0.25, pure_id=-1)                                # neither pure nor target states
    with pytest.raises(qb.QBError):
        eng.set_objective(qb.OBJ_JTRACE, np.ones(4) / 4, 0.25, pure_id=4)                                 # |4> does not exist for N = 4
    eng.set_objective(qb.OBJ_JTRACE, np.ones(4) / 4, 0.25, pure_id=1)
    with pytest.raises(qb.QBError):
        eng.objective(4)                                                   # no forward sweep yet
    eng.close()


# ---- Schroedinger solver on large Hilbert spaces, 256 < N <= 4096 (large-state kernels, qb_bigs_q*.cu) -----------------
def _schroedinger_big(levels, ntime, linsolve, flags, jkl=0.002, seed=5):
    Q = len(levels)
    npairs = Q * (Q - 1) // 2
    rng = np.random.default_rng(seed)
    T = 0.04 * ntime
    return qb.make_desc(
        nlevels=levels, lindblad="none", detuning=list(TWO_PI * rng.uniform(-0.05, 0.05, Q)), selfkerr=list(TWO_PI * rng.uniform(0.01, 0.03, Q)),
        crosskerr=list(TWO_PI * rng.uniform(0.0, 0.002, npairs)), Jkl=[TWO_PI * jkl] * npairs, eta=list(TWO_PI * rng.uniform(-0.3, 0.3, npairs)),
        ntime=ntime, total_time=T, linsolve=linsolve, maxiter=30,
        segments=[[{"type": "spline", "nsplines": 6, "tstart": 0.0, "tstop": T}] for _ in range(Q)],
        carriers=[[0.0, TWO_PI * 0.05] for _ in range(Q)], flags=flags)


SBIG = [((20, 20), 8, 0.002), ((8, 8, 8), 8, 0.002), ((4, 4, 4, 4, 4), 6, 0.0), ((12, 12, 12), 5, 0.002), ((16, 16, 16), 4, 0.0), ((3000,), 4, 0.0), ((30, 9), 6, 0.002)]


@pytest.mark.parametrize("flags", [qb.FLAG_STORE_STAGE, qb.FLAG_RECOMPUTE_STAGE, 0])
@pytest.mark.parametrize("linsolve", ["gmres", "neumann"])
@pytest.mark.parametrize("levels,nt,jkl", SBIG)
def test_big_schroedinger_matches_oracle(oracle, levels, nt, jkl, linsolve, flags):
    if flags == 0 and linsolve == "neumann":
        pytest.skip("covered by the gmres variant")
    if max(levels) > 31:
        with pytest.raises(qb.QBError):
            Engine(_schroedinger_big(levels, nt, linsolve, flags, jkl))     # more than 31 levels per oscillator: loud, not wrong
        return
    if linsolve == "neumann" and max(levels) >= 20 and flags != qb.FLAG_RECOMPUTE_STAGE:
        # ||dt/2 M|| > 1 at 20+ levels: the Neumann series stops at maxiter, and a stored stage equals the reference's recomputed one
        # only up to the linear-solve residual; the recompute mode (the reference's exact order of operations) covers this regime
        pytest.skip("non-convergent Neumann regime: recompute mode only")
    desc = _schroedinger_big(levels, nt, linsolve, flags, jkl)
    N = int(np.prod(levels))
    nd = desc_ndesign(desc)
    rng = np.random.default_rng(11)
    alpha = TWO_PI * 0.003 * (2.0 * rng.random(nd) - 1.0)
    psi0 = np.zeros((3, 2 * N))
    psi0[0, 0] = 1.0
    psi0[1, N // 3] = 1.0
    v = rng.standard_normal(2 * N)
    psi0[2] = v / np.linalg.norm(v)
    rbar = rng.standard_normal(psi0.shape)
    eng = Engine(desc)
    eng.set_params(alpha)
    for t in (0.0, 0.6 * desc.total_time):
        for tr in (False, True):
            assert relerr(eng.rhs_apply(t, psi0, transpose=tr), oracle.rhs_apply(desc, alpha, t, psi0, transpose=tr)) < 1e-13
    psiT = eng.forward(psi0, store=True)
    st = eng.stats()
    mid = eng.get_state(2, nt // 2)
    grad, xb0 = eng.adjoint(rbar, want_rho0_bar=True)
    eng.close()
    gsum = np.zeros(nd)
    iters = 0
    for i in range(3):
        pTo, traj, it, _ = oracle.solve_ode(desc, alpha, psi0[i], want_traj=True)
        iters += it
        assert relerr(psiT[i], pTo) < 1e-11, (levels, i, relerr(psiT[i], pTo))
        if linsolve == "gmres":      # the implicit midpoint rule preserves the norm up to the solver tolerance (Neumann may stop at maxiter: ||dt/2 M|| > 1 at 20+ levels)
            assert abs(np.linalg.norm(psiT[i]) - 1.0) < 1e-9
        if i == 2:
            assert relerr(mid, traj[nt // 2]) < 1e-11
        go, xbo, _, _ = oracle.solve_adjoint(desc, alpha, rbar[i], pTo, None)
        gsum += go
        assert relerr(xb0[i], xbo) < 1e-9, (levels, i, relerr(xb0[i], xbo))
    assert st["linsolve_iters"] == iters
    assert relerr(grad, gsum) < (5e-9 if flags != qb.FLAG_RECOMPUTE_STAGE else 1e-9), relerr(grad, gsum)
