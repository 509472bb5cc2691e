"""CPU checks of bench.py: the flop model of SURVEY section 8d, and the reference arm's JSON contract (runs the unmodified
reference binary oracle/_ref/quandary -- or the C port -- on a 5-step sample)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_flop_model_matches_survey_table():
    # F_rhs = 72 (C1), 640 (C2), 3648 (C3), 18944 (C4), 568320 (C5 with three couplings)
    assert bench.flops_per_rhs([2, 2], False, [0.0]) == 72
    assert bench.flops_per_rhs([2, 2], True, [0.0]) == 640
    assert bench.flops_per_rhs([2, 2, 2], True, [0.0] * 3) == 3648
    assert bench.flops_per_rhs([2, 2, 2, 2], True, [0.0] * 6) == 18944
    assert bench.flops_per_rhs([4, 4, 4], True, [0.001] * 3) == 568320


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-ntime", "5"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "steps/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("C4") and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_at_n_gpus_times_the_same_job_as_the_b200_arm():
    """N > 1: the b200 arm's headline is BASELINE config 4 as written (256 initial conditions divided over the ranks, strong scaling);
    the reference arm (rank 0 only) times the same 256 initial conditions, so the driver compares equal configurations."""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-ntime", "5"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][-1])
    assert d["scaling"] == "strong" and d["n_gpus"] == 2
    assert d["config"]["ninit_total"] == 256 and d["config"]["ninit_per_gpu"] == 128

