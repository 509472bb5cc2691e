#!/usr/bin/env python
"""Development aid: per-phase cycle counts of one forward step of the C4 qubit kernel (E = 2), from a library built with
-DQB_PHASE_CLOCKS (scripts/_clk/libquandary_b200.so; see the QB_CLK marks in csrc/qb_qubit.cuh)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import quandary_b200 as qb
from quandary_b200.systems import cfg_system, random_alpha, basis_states
qb._lib = qb.load_library(os.path.join(ROOT, "scripts", "_clk", "libquandary_b200.so"))
os.environ["QB_QUBIT_E"] = "2"
for n in (32, 256):
    desc, ninit = cfg_system("C4", ntime=2000, flags=qb.FLAG_STORE_STAGE)
    eng = qb.Engine(desc)
    eng.set_params(random_alpha(qb.desc_ndesign(desc), amp=0.005))
    rho0 = basis_states(16, True)[:n]
    eng.upload_states(rho0, 0)
    print("ninit", n, flush=True)
    for _ in range(2):
        eng.forward_dev(n, store=2)
        eng.sync()
    print(eng.stats()["last_fwd_ms"], "ms for 2000 steps")
    eng.close()
