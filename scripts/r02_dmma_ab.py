#!/usr/bin/env python
"""A/B of the operator formulation at N = 64 (BASELINE config 5): stencil kernel (qbig::bapply_kernel) against the dense tensor-core
kernel (qdense::dense_apply_kernel, DMMA), same states, same control values; per-application time from launches with R and 1
applications inside (set-up and state staging cancel).  Prints one JSON object."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import quandary_b200 as qb  # noqa: E402
from quandary_b200.systems import cfg_system, random_alpha  # noqa: E402

desc, _ = cfg_system("C5", ntime=8)
nd = qb.desc_ndesign(desc)
eng = qb.Engine(desc)
eng.set_params(random_alpha(nd, amp=0.05))
nb = 296
x = np.random.default_rng(1).standard_normal((nb, 2 * qb.desc_dim(desc)))
eng.upload_states(x, 0)
eng.upload_states(x, 1)
t = 0.37 * desc.total_time
out = {"system": "C5 (4,4,4), N=64, dim 4096, 3 Jkl couplings", "states": nb}
R = 41
for name, dense in (("stencil", 0), ("dmma", 1)):
    for _ in range(2):
        eng.apply_bench(t, nb, R, dense)
    tR = min(eng.apply_bench(t, nb, R, dense) for _ in range(5))
    t1 = min(eng.apply_bench(t, nb, 1, dense) for _ in range(5))
    per = (tR - t1) / (R - 1)        # ms per application of all nb states
    flops = {"stencil": 568320.0, "dmma": 8 * 2 * 64 ** 3 + 0.0}[name] * nb
    out[name] = {"ms_launch_R%d" % R: tR, "ms_launch_R1": t1, "us_per_application_per_state_slot": 1e3 * per / max(1, nb / 148.0) / 1.0,
                 "ms_per_application_all_states": per, "tflops_nominal": flops / (per * 1e-3) / 1e12}
out["dmma_peak_tflops"] = qb.measure_fp64_mma_peak(0)
out["dfma_peak_tflops"] = qb.measure_fp64_peak(0)
out["speed_ratio_stencil_over_dmma"] = out["dmma"]["ms_per_application_all_states"] / out["stencil"]["ms_per_application_all_states"]
eng.close()
print(json.dumps(out))
