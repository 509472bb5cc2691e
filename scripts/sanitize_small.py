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
    print(name, ls, flags, float(np.abs(rT).sum()), float(np.abs(g).sum()))
    e.close()
