"""The reference's OWN regression cases (tests/golden/ref_regression/, copied verbatim from /root/reference/tests/regression by
tests/golden/fetch_ref_regression.py: its .cfg files and its golden base/ outputs) run through quandary_b200_main and compared
exactly as the reference's suite does (tests/regression/regression_test.py:17-18,85-88: rtol 1e-7, atol 1e-15)."""
import glob
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "quandary_b200", "quandary_b200_main")
GOLD = os.path.join(ROOT, "tests", "golden", "ref_regression")
RTOL, ATOL = 1e-7, 1e-15
CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "*")))
TAO_CASES = {"cnot", "xgate"}      # runtype = optimization: only the evaluation at the start point is optimizer independent


def load(path):
    rows = []
    with open(path) as f:
        for line in f:
            if line.startswith("#") or not line.strip():
                continue
            rows.append([float(t) for t in line.split()])
    return np.array(rows)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_reference_regression_case(case, tmp_path):
    src = os.path.join(GOLD, case)
    for f in os.listdir(src):
        if f != "base":
            shutil.copy(os.path.join(src, f), tmp_path)
    cfg = case + ".cfg"
    if case in TAO_CASES:      # evaluate the start point only
        text = open(os.path.join(tmp_path, cfg)).read() + "\noptim_maxiter = 0\n"
        open(os.path.join(tmp_path, cfg), "w").write(text)
    r = subprocess.run([BIN, cfg, "--quiet"], cwd=tmp_path, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    golden = sorted(glob.glob(os.path.join(src, "base", "*.dat")))
    assert golden
    for g in golden:
        fn = os.path.basename(g)
        out = os.path.join(tmp_path, "data_out", fn)
        assert os.path.exists(out), f"{fn} not written"
        a, b = load(out), load(g)
        if case in TAO_CASES:
            keep = [0, 1, 4, 5, 6, 7, 8, 9, 10]     # drop ||Pr(grad)|| and LS step (TAO internals)
            a, b = a[:1, keep], b[:1, keep]
        assert a.shape == b.shape, (fn, a.shape, b.shape)
        bad = np.abs(a - b) > ATOL + RTOL * np.abs(b)
        assert not bad.any(), (case, fn, int(bad.sum()), a[bad][:3], b[bad][:3])


@pytest.mark.gpu
def test_reference_cnot_optimization_converges(tmp_path):
    """runtype = optimization on the reference's cnot.cfg: the projected L-BFGS (not TAO) drives the same objective to the
    same stopping criterion (optim_inftol 1e-5 on the infidelity); TAO needs 17 iterations, anything comparable is fine."""
    src = os.path.join(GOLD, "cnot")
    shutil.copy(os.path.join(src, "cnot.cfg"), tmp_path)
    r = subprocess.run([BIN, "cnot.cfg", "--quiet"], cwd=tmp_path, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    hist = load(os.path.join(tmp_path, "data_out", "optim_history.dat"))
    ref = load(os.path.join(src, "base", "optim_history.dat"))
    assert abs(hist[0, 1] - ref[0, 1]) <= 1e-7 * abs(ref[0, 1])          # same start
    assert np.all(np.diff(hist[:, 1]) <= 1e-12)                           # monotone objective
    assert hist[-1, 4] > 1.0 - 1e-4 and hist.shape[0] <= 60, (hist[-1], hist.shape)
    for k in range(2):                                                    # the final iterate is simulated with outputs on
        assert os.path.exists(os.path.join(tmp_path, "data_out", f"expected{k}.iinit0000.dat"))
