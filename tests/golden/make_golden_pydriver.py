#!/usr/bin/env python
"""Generates tests/golden/pydriver/<case>/: configuration files written by the reference's OWN Python driver
(/root/reference/quandary.py, class Quandary, its private __dump: config.cfg + targetgate.dat exactly as a user of the Python
interface produces them) and the output files the UNMODIFIED reference binary (oracle/_ref/quandary) writes for them.
tests/test_host_gpu.py runs `quandary_b200_main config.cfg` on the same files and compares every output file.

The only post-edit of the driver's config: output_frequency (the driver writes 1, i.e. every time step; the fixtures keep every
100th step to stay small) and optim_maxiter for the optimisation case.  matplotlib is absent in this image and is stubbed; the
driver only needs it for plotting.

    python tests/golden/make_golden_pydriver.py
"""
import os
import re
import shutil
import subprocess
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "quandary")
OUT = os.path.join(HERE, "pydriver")

mpl = types.ModuleType("matplotlib")
plt = types.ModuleType("matplotlib.pyplot")
mpl.pyplot = plt
sys.modules["matplotlib"] = mpl
sys.modules["matplotlib.pyplot"] = plt
sys.path.insert(0, "/root/reference")
import numpy as np  # noqa: E402
import quandary as refdriver  # noqa: E402

CNOT = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=complex)
XGATE = np.array([[0, 1], [1, 0]], dtype=complex)


def cases():
    yield "cnot_two_transmons_guard_level_schroedinger", "gradient", refdriver.Quandary(
        Ne=[2, 2], Ng=[1, 1], freq01=[4.10595, 4.81526], selfkerr=[0.2198, 0.2252], crosskerr=[0.01], T=20.0, targetgate=CNOT,
        maxctrl_MHz=[10.0, 10.0], rand_seed=4321, verbose=False, nsplines=8)
    yield "xgate_transmon_guard_level_lindblad", "gradient", refdriver.Quandary(
        Ne=[2], Ng=[1], freq01=[4.10595], selfkerr=[0.2198], T=15.0, targetgate=XGATE, T1=[80.0], T2=[40.0], maxctrl_MHz=[8.0], rand_seed=77,
        verbose=False, nsplines=10, costfunction="Jfrobenius", initialcondition="basis")
    yield "cnot_two_transmons_optimization_row0", "optimization", refdriver.Quandary(
        Ne=[2, 2], Ng=[1, 1], freq01=[4.10595, 4.81526], selfkerr=[0.2198, 0.2252], crosskerr=[0.01], T=20.0, targetgate=CNOT,
        maxctrl_MHz=[10.0, 10.0], rand_seed=99, verbose=False, nsplines=8, control_enforce_BC=True)


def more_cases():
    # state preparation |0> -> (|0> + |2>)/sqrt(2) on a 3-level system with one guard level: target STATE from file, pure initial condition
    t = np.array([1.0, 0.0, 1.0, 0.0]) / np.sqrt(2.0)
    yield "state_preparation_lindblad_target_file", "gradient", refdriver.Quandary(
        Ne=[3], Ng=[1], freq01=[4.10595], selfkerr=[0.2198], T=12.0, targetstate=list(t), T1=[60.0], T2=[30.0], maxctrl_MHz=[8.0], rand_seed=5,
        verbose=False, nsplines=10, initialcondition="pure, 0", costfunction="Jtrace")
    yield "state_preparation_schroedinger_target_file", "gradient", refdriver.Quandary(
        Ne=[2, 2], Ng=[1, 1], freq01=[4.10595, 4.81526], selfkerr=[0.2198, 0.2252], crosskerr=[0.01], T=10.0,
        targetstate=list(np.array([1.0, 0.0, 0.0, 1.0]) / np.sqrt(2.0)), maxctrl_MHz=[10.0, 10.0], rand_seed=6, verbose=False, nsplines=8,
        initialcondition="pure, 0, 0", costfunction="Jfrobenius")


def main():
    only = sys.argv[1:]
    if not only and os.path.exists(OUT):
        shutil.rmtree(OUT)
    import itertools
    for name, runtype, q in itertools.chain(cases(), more_cases()):
        if only and name not in only:
            continue
        if os.path.exists(os.path.join(OUT, name)):
            shutil.rmtree(os.path.join(OUT, name))
        d = os.path.join(OUT, name)
        os.makedirs(d)
        q._Quandary__dump(pcof0=[], runtype=runtype, datadir=d)
        cfg = open(os.path.join(d, "config.cfg")).read()
        cfg = re.sub(r"output_frequency = 1\n", "output_frequency = 100\n", cfg)
        if runtype == "optimization":
            cfg = re.sub(r"optim_maxiter= \d+\n", "optim_maxiter= 0\n", cfg)
        open(os.path.join(d, "config.cfg"), "w").write(cfg)
        inputs = set(os.listdir(d))
        r = subprocess.run([REF, "config.cfg", "--quiet"], cwd=d, env=dict(os.environ, QSHIM_NP="1"), capture_output=True, text=True)
        assert r.returncode == 0, (name, r.stdout[-400:], r.stderr[-400:])
        base = os.path.join(d, "base")
        os.makedirs(base)
        for f in sorted(set(os.listdir(d)) - inputs - {"base"}):       # datadir = ./ : outputs land next to the config
            if f in ("config_log.dat", "timing.dat"):
                os.remove(os.path.join(d, f))
            else:
                os.rename(os.path.join(d, f), os.path.join(base, f))
        print(name, len(os.listdir(base)), "files")


if __name__ == "__main__":
    main()
