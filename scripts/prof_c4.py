import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import qb_systems
from quandary_b200 import Engine, desc_ndesign
ntime = int(sys.argv[1]); ls = sys.argv[2]
d, ninit = qb_systems.cfg_system("C4", ntime=ntime, linsolve=ls)
a = qb_systems.random_alpha(desc_ndesign(d), amp=0.005)
rho0 = qb_systems.basis_states(16, True)
e = Engine(d); e.set_params(a)
rbar = np.random.default_rng(0).standard_normal(rho0.shape) / ninit
e.upload_states(rho0); e.forward_dev(ninit, store=True); e.sync()
e.upload_states(rbar, 1); e.adjoint_dev(ninit); e.sync()
print(e.stats())
