import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import qb_oracle, qb_systems
from quandary_b200 import Engine, desc_ndesign
o = qb_oracle.load()
for name in ["C1", "C2", "G5"]:
    for ls in ["gmres", "neumann"]:
        for tol in [1e-10, 1e-14]:
            for nt in [1, 2, 60]:
                d, ni = qb_systems.cfg_system(name, ntime=nt, linsolve=ls)
                d.linsolve_abstol = tol
                a = qb_systems.random_alpha(desc_ndesign(d), amp=0.02)
                N = int(np.prod([d.nlevels[k] for k in range(d.nosc)]))
                r0 = qb_systems.basis_states(N, d.lindblad != 0)[:1]
                e = Engine(d); e.set_params(a)
                x = e.forward(r0, store=True)
                st = e.stats()
                xo, traj, it, ev = o.solve_ode(d, a, r0[0], want_traj=True)
                print(name, ls, tol, nt, "iters", st["linsolve_iters"], it, "evals", st["rhs_evals"], ev, "err", np.max(np.abs(x[0] - xo)))
                e.close()
