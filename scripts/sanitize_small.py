"""Tiny forward+adjoint runs of every kernel family, for compute-sanitizer (racecheck / memcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import quandary_b200 as qb
from quandary_b200.systems import basis_states, cfg_system, random_alpha

for name, nt, flags, ls in [("C3", 4, qb.FLAG_STORE_STAGE, "gmres"), ("C4", 3, qb.FLAG_RECOMPUTE_STAGE, "gmres"), ("C4", 3, qb.FLAG_STORE_STAGE, "neumann"),
                            ("C2", 4, 0, "gmres"), ("C1", 4, 0, "gmres"), ("G32J", 4, qb.FLAG_RECOMPUTE_STAGE, "gmres"), ("G333", 2, qb.FLAG_STORE_STAGE, "gmres"),
                            ("G333", 2, qb.FLAG_RECOMPUTE_STAGE, "neumann")]:
    desc, ninit = cfg_system(name, ntime=nt, linsolve=ls, flags=flags)
    N = int(np.prod([desc.nlevels[k] for k in range(desc.nosc)]))
    rho0 = basis_states(N, desc.lindblad != 0)[:3]
    rbar = np.random.default_rng(0).standard_normal(rho0.shape)
    e = qb.Engine(desc)
    e.set_params(random_alpha(qb.desc_ndesign(desc), amp=0.02))
    rT = e.forward(rho0, store=True)
    g = e.adjoint(rbar)
    # final-time objective kernels on the same final states
    w = np.full(3, 1.0 / 3)
    e.set_objective(qb.OBJ_JTRACE, w, 1.0 / 3, pure_id=-1, target=rbar, purity=np.ones(3))
    o = e.objective(3)
    e.objective_bar(3, -1.0, 0.5 if desc.lindblad == 0 else 0.0)
    e.adjoint_dev(3)
    g2 = e.download_grad()
    print(name, ls, flags, float(np.abs(rT).sum()), float(np.abs(g).sum()), float(np.abs(o).sum()), float(np.abs(g2).sum()))
    e.close()

# large-state Schroedinger kernels (N = 400), with the dpdm penalty
T = 0.08
desc = qb.make_desc(nlevels=[20, 20], lindblad="none", detuning=[0.1, -0.2], selfkerr=[0.05, 0.04], crosskerr=[0.01], Jkl=[0.02], eta=[0.7], ntime=3, total_time=T,
                    linsolve="gmres", maxiter=12, segments=[[{"type": "spline", "nsplines": 5, "tstart": 0.0, "tstop": T}]] * 2, carriers=[[0.0, 0.3]] * 2,
                    flags=qb.FLAG_RECOMPUTE_STAGE)
rng = np.random.default_rng(3)
psi0 = rng.standard_normal((2, 800)) / 20.0
e = qb.Engine(desc)
e.set_params(random_alpha(qb.desc_ndesign(desc), amp=0.02))
jb = np.array([0.5, 0.25])
C = qb.C
e._check(e.lib.qb_engine_set_dpdm(e.h, jb.ctypes.data_as(C.POINTER(C.c_double)), 2))
pT = e.forward(psi0, store=True)
g = e.adjoint(rng.standard_normal(psi0.shape))
print("S2020 dpdm", float(np.abs(pT).sum()), float(np.abs(g).sum()))
e.close()
