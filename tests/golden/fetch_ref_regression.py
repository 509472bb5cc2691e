#!/usr/bin/env python
"""Copies the in-scope cases of the reference's OWN regression suite (/root/reference/tests/regression: .cfg, auxiliary input
files and the golden `base/` output files) into tests/golden/ref_regression/ so that tests/test_ref_regression_gpu.py can run
them through quandary_b200_main on the GPU box, where /root/reference does not exist.  Only data files are copied (test inputs
and expected outputs), no reference source code.  Large full-state files are skipped (listed below).

In scope = matrix-free-expressible Lindblad/Schroedinger runs without TAO and without Hamiltonian files:
  AxC, AxC_grad_initBasis0, AxC_grad_schroedinger, AxC_initDiag0, AxC_initEnsemble, AxC_initFile, pipulse, xgate_sparsemat and nlevels_4_4_4_4 (usematfree = false in the .cfg; run matrix-free, same mathematics),
  cnot and xgate (runtype optimization: row 0 of optim_history.dat, the evaluation at the start point, is TAO independent).
Out of scope: state-to-state_spline0 (TAO iterates), spinchain_N8 (8 oscillators),
hamiltonian-reader* (Hamiltonians from files: sparse-matrix-only feature), AxC_grad (no golden files in the reference tree).
"""
import glob
import os
import shutil

REF = "/root/reference/tests/regression"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ref_regression")
CASES = {
    "AxC": ["optim_history.dat", "population*.dat", "expected*.dat"],                    # rho*.dat: 3.3 MB, skipped
    "AxC_grad_initBasis0": ["grad.dat", "expected*.dat", "optim_history.dat"],
    "AxC_initDiag0": ["rho*.dat", "optim_history.dat"],
    "AxC_initEnsemble": ["rho*.dat", "optim_history.dat"],
    "AxC_initFile": ["rho*.dat", "optim_history.dat"],
    "pipulse": ["optim_history.dat", "expected*.dat", "population*.iinit0000.dat"],      # rho*.dat / other populations: 17 MB, skipped
    "xgate_sparsemat": ["grad.dat", "optim_history.dat", "rho*.dat", "population*.dat"],
    "AxC_grad_schroedinger": ["grad.dat", "optim_history.dat", "rho*.dat"],          # dpdm + energy + weighted-J penalties, Schroedinger
    "nlevels_4_4_4_4": ["expected*.dat", "population[0-3]*.dat"],                     # population_composite: 2.6 MB, skipped; run matrix-free
    "cnot": ["optim_history.dat"],
    "xgate": ["optim_history.dat"],
}


def main():
    if os.path.exists(OUT):
        shutil.rmtree(OUT)
    for name, patterns in CASES.items():
        src, dst = os.path.join(REF, name), os.path.join(OUT, name)
        os.makedirs(os.path.join(dst, "base"))
        for f in os.listdir(src):
            p = os.path.join(src, f)
            if os.path.isfile(p):
                shutil.copy(p, dst)
        n = 0
        for pat in patterns:
            for g in glob.glob(os.path.join(src, "base", pat)):
                shutil.copy(g, os.path.join(dst, "base"))
                n += 1
        print(name, n, "golden files")


if __name__ == "__main__":
    main()
