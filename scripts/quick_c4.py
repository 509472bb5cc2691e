"""Quick timing of BASELINE config 4 (N=16 Lindblad, 256 ICs) on one GPU: device-resident forward+adjoint."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import qb_systems
from quandary_b200 import Engine, desc_ndesign
ntime = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
for ls in ["gmres", "neumann"]:
    d, ninit = qb_systems.cfg_system("C4", ntime=ntime, linsolve=ls)
    a = qb_systems.random_alpha(desc_ndesign(d), amp=0.005)
    rho0 = qb_systems.basis_states(16, True)
    e = Engine(d); e.set_params(a)
    rng = np.random.default_rng(0)
    rbar = rng.standard_normal(rho0.shape) / ninit
    for rep in range(3):
        e.upload_states(rho0); e.forward_dev(ninit, store=True); e.sync()
        t0 = time.time()
        e.upload_states(rho0); e.forward_dev(ninit, store=True); e.sync()
        t1 = time.time()
        e.upload_states(rbar, 1); e.adjoint_dev(ninit); e.sync()
        t2 = time.time()
        st = e.stats()
        steps = ninit * ntime
        print(ls, "rep", rep, "fwd %.1f ms (kernel %.1f) adj %.1f ms (kernel %.1f)" % ((t1 - t0) * 1e3, st["last_fwd_ms"], (t2 - t1) * 1e3, st["last_adj_ms"]),
              "steps/s %.3e" % (steps / (t2 - t0)), "rhs_evals/s %.3e" % (st["rhs_evals"] / (t2 - t0)), "iters/step %.2f" % (st["linsolve_iters"] / (2 * steps)),
              "evals", st["rhs_evals"])
    e.close()
