#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libquandary_ref.so = /root/reference/src
compiled against the sequential PETSc/MPI shim, driven through oracle/ref_hooks.cpp).  Run in the build container
(needs /root/reference for `make -C oracle ref`); the fixtures travel, the reference does not.

    python tests/golden/make_golden.py

Per system (quandary_b200.systems, reduced ntime) the fixture holds, all produced by reference code:
  alpha                      design vector fed to both sides
  ctrl_t, ctrl_p, ctrl_q     Oscillator::evalControl                          src/oscillator.cpp:281-337
  x, rhs_t, y, yT            MasterEq::assemble_RHS + MatMult/MatMultTranspose src/mastereq.cpp:657-739
  dr_x, dr_xbar, dr_t, dr_g  MasterEq::compute_dRHS_dParams                    src/mastereq.cpp:970-1276
  ic_idx, rho0               OptimTarget::prepareInitialState (basis)          src/optimtarget.cpp:450-698
  rhoT, traj_n, traj         TimeStepper::solveODE (+ stored states)           src/timestepper.cpp:96-181
  rbar, redgrad              TimeStepper::solveAdjointODE                      src/timestepper.cpp:184-253
  obj, fid, grad             OptimProblem::evalGradF (gate systems only)       src/optimproblem.cpp:342-538
for linearsolver_type = gmres and neumann.  Level tuples without a matrix-free instantiation in the reference
(src/mastereq.cpp:2977-3239: G32J, G232S) go through its sparse-matrix operator (usematfree = false).
"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from quandary_b200 import desc_dim, desc_ndesign  # noqa: E402
from quandary_b200.systems import cfg_system, cfg_text, random_alpha, ref_has_matfree, system_params  # noqa: E402

LIB = os.path.join(ROOT, "oracle", "_ref", "libquandary_ref.so")
dp = C.POINTER(C.c_double)


def P(a):
    return a.ctypes.data_as(dp)


def load():
    lib = C.CDLL(LIB)
    lib.qref_create.argtypes = [C.c_char_p, C.c_int]; lib.qref_create.restype = C.c_void_p
    for name in ("qref_dim", "qref_ndesign", "qref_ninit", "qref_ntime"):
        getattr(lib, name).argtypes = [C.c_void_p]; getattr(lib, name).restype = C.c_int
    lib.qref_set_params.argtypes = [C.c_void_p, dp]
    lib.qref_eval_control.argtypes = [C.c_void_p, C.c_int, C.c_double, dp, dp]
    lib.qref_rhs_apply.argtypes = [C.c_void_p, C.c_double, C.c_int, dp, dp]
    lib.qref_drhs_dparams.argtypes = [C.c_void_p, C.c_double, dp, dp, C.c_double, dp]
    lib.qref_initial_state.argtypes = [C.c_void_p, C.c_int, dp]
    lib.qref_solve_ode.argtypes = [C.c_void_p, dp, dp, dp]
    lib.qref_solve_adjoint.argtypes = [C.c_void_p, dp, dp, dp, dp]
    lib.qref_eval_grad_f.argtypes = [C.c_void_p, dp, dp, dp]; lib.qref_eval_grad_f.restype = C.c_double
    return lib


SYSTEMS = {"C1": 40, "C2": 40, "C3": 30, "C4": 12, "G32J": 40, "G44J": 16, "G232S": 40, "G5": 40, "Q5": 30,
           "G333": 8, "C5": 5, "A320": 6, "Q5L": 8}   # the last four: large-state kernels (dim 729, 4096, 3600, 1024)


def make(lib, name, ntime, linsolve, workdir):
    desc, ninit = cfg_system(name, ntime=ntime, linsolve=linsolve)
    nd, dim = desc_ndesign(desc), desc_dim(desc)
    alpha = random_alpha(nd, amp=0.02)
    pfile = os.path.join(workdir, f"params_{name}.dat")
    with open(pfile, "w") as f:
        f.write("\n".join(repr(float(v)) for v in alpha) + "\n")
    cfg = os.path.join(workdir, f"{name}_{linsolve}.cfg")
    with open(cfg, "w") as f:
        f.write(cfg_text(name, ntime=ntime, linsolve=linsolve, params_file=pfile, datadir=os.path.join(workdir, "data_out"),
                         usematfree=ref_has_matfree(name)))
    h = lib.qref_create(cfg.encode(), 1 if desc.lindblad != 0 else 0)
    assert lib.qref_dim(h) == dim and lib.qref_ndesign(h) == nd and lib.qref_ninit(h) == ninit, (name, lib.qref_dim(h), lib.qref_ndesign(h), lib.qref_ninit(h))
    lib.qref_set_params(h, P(alpha))
    out = {"alpha": alpha, "ntime": np.array(ntime)}
    rng = np.random.default_rng(4321)
    # controls
    T = desc.total_time
    ct = np.concatenate([np.linspace(0.0, T, 9), rng.random(4) * T])
    cp, cq = np.zeros((ct.size, desc.nosc)), np.zeros((ct.size, desc.nosc))
    for i, t in enumerate(ct):
        for k in range(desc.nosc):
            a, b = C.c_double(), C.c_double()
            lib.qref_eval_control(h, k, float(t), C.byref(a), C.byref(b))
            cp[i, k], cq[i, k] = a.value, b.value
    out.update(ctrl_t=ct, ctrl_p=cp, ctrl_q=cq)
    # operator
    x = rng.standard_normal((2, 2 * dim))
    ts = np.array([0.0, 0.37 * T, T])
    y = np.zeros((ts.size, 2, 2 * dim)); yT = np.zeros_like(y)
    for i, t in enumerate(ts):
        for j in range(2):
            lib.qref_rhs_apply(h, float(t), 0, P(x[j]), P(y[i, j]))
            lib.qref_rhs_apply(h, float(t), 1, P(x[j]), P(yT[i, j]))
    out.update(x=x, rhs_t=ts, y=y, yT=yT)
    drx, drb = rng.standard_normal(2 * dim), rng.standard_normal(2 * dim)
    drg = np.zeros(max(1, nd))
    lib.qref_drhs_dparams(h, float(0.41 * T), P(drx), P(drb), 0.7, P(drg))
    out.update(dr_x=drx, dr_xbar=drb, dr_t=np.array(0.41 * T), dr_g=drg[:nd])
    # sweeps
    idx = np.unique(np.linspace(0, ninit - 1, 3).astype(int))
    rho0 = np.zeros((idx.size, 2 * dim)); rhoT = np.zeros_like(rho0)
    rbar = rng.standard_normal(rho0.shape)
    red = np.zeros((idx.size, max(1, nd)))
    traj_n = np.array([0, 1, ntime // 2, ntime])
    traj = np.zeros((idx.size, traj_n.size, 2 * dim))
    full = np.zeros((ntime + 1, 2 * dim))
    for i, ii in enumerate(idx):
        lib.qref_initial_state(h, int(ii), P(rho0[i]))
        lib.qref_solve_ode(h, P(rho0[i]), P(rhoT[i]), P(full) if desc.lindblad != 0 else None)
        if desc.lindblad != 0:
            traj[i] = full[traj_n]
        lib.qref_solve_adjoint(h, P(rbar[i]), P(rhoT[i]), P(red[i]), None)
    out.update(ic_idx=idx, rho0=rho0, rhoT=rhoT, rbar=rbar, redgrad=red[:, :nd], traj_n=traj_n, traj=traj)
    if system_params(name)["gate"]:
        g = np.zeros(nd); fid = C.c_double()
        obj = lib.qref_eval_grad_f(h, P(alpha), P(g), C.byref(fid))
        out.update(obj=np.array(obj), fid=np.array(fid.value), grad=g)
    return out


def main():
    lib = load()
    wd = tempfile.mkdtemp(prefix="qb_golden_")
    for name, ntime in SYSTEMS.items():
        for ls in ("gmres", "neumann"):
            data = make(lib, name, ntime, ls, wd)
            np.savez_compressed(os.path.join(HERE, f"{name}_{ls}.npz"), **data)
            print(name, ls, "objective" if "obj" in data else "", float(data["obj"]) if "obj" in data else "")


if __name__ == "__main__":
    main()
