#!/usr/bin/env python
"""Generates tests/golden/host/<case>/: a reference .cfg (+ params.dat) and the output files the UNMODIFIED reference binary
(oracle/_ref/quandary, reference sources + PETSc shim) writes for it.  tests/test_host_gpu.py runs the B200 entry point
`quandary_b200_main` on the same .cfg and compares every file with the reference regression suite's own tolerance
(rtol 1e-7, atol 1e-15; /root/reference/tests/regression/regression_test.py:17-18).

    python tests/golden/make_golden_host.py
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from quandary_b200 import desc_ndesign  # noqa: E402
from quandary_b200.systems import cfg_system, cfg_text, random_alpha  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "quandary")
OUT = os.path.join(HERE, "host")

CASES = {
    # name: (system, kwargs for cfg_text, write params.dat?)
    "C2_simulation_outputs": ("C2", dict(ntime=100, runtype="simulation", outputs="expectedEnergy, population, fullstate", output_frequency=10), True),
    "C3_gradient_random_init": ("C3", dict(ntime=60, runtype="gradient", control_init="random, 0.004"), False),
    "C1_gradient_schroedinger": ("C1", dict(ntime=100, runtype="gradient", control_init="constant, 0.005", outputs="expectedEnergy", output_frequency=25), False),
    "C4_gradient_neumann": ("C4", dict(ntime=20, runtype="gradient", linsolve="neumann"), True),
    "G33_gradient_frobenius_pure_diagonal": ("G33", dict(ntime=50, runtype="gradient", objective="Jfrobenius", target="pure, 1, 0", initialcondition="diagonal",
                                                      weights="1.0, 2.0, 0.5"), True),
    "G33_simulation_measure_3states": ("G33", dict(ntime=50, runtype="simulation", objective="Jmeasure", target="pure, 0, 2", initialcondition="3states",
                                                  outputs="population, expectedEnergyComposite", output_frequency=5), True),
    "C2_gradient_frobenius_swap": ("C2", dict(ntime=80, runtype="gradient", objective="Jfrobenius", target="gate, swap", control_init="random, 0.003"), False),
    "A320_gradient_penalties_basis0": ("A320", dict(ntime=60, runtype="gradient", objective="Jmeasure", target="pure, 0, 0", initialcondition="basis, 0",
                                                    weights="1.0, 1.0", penalty=1.0, penalty_param=0.5, penalty_energy=0.1, enforceBC=True,
                                                    outputs="expectedEnergy", output_frequency=20), True),
    "G33_gradient_guard_levels_cnot": ("G33", dict(ntime=50, runtype="gradient", objective="Jtrace", target="gate, cnot", nessential=[2, 2], penalty=1.0,
                                                    penalty_param=0.8, penalty_energy=0.05, control_init="random, 0.003"), False),
    "C2_gradient_penalty_frobenius_swap": ("C2", dict(ntime=80, runtype="gradient", objective="Jfrobenius", target="gate, swap", penalty=0.7, penalty_param=2.0,
                                                       control_init="random, 0.003"), False),
    "G3_simulation_pipulse_leakage": ("G3", dict(ntime=200, runtype="simulation", objective="Jmeasure", target="pure, 0", nessential=[2], penalty=1.0,
                                                  penalty_param=0.5, enforceBC=True, pipulse="0, 0.2, 0.9, 1.3", control_init="constant, 0.004",
                                                  outputs="expectedEnergy, population, fullstate", output_frequency=20), False),
    "C1_gradient_penalty_schroedinger_measure": ("C1", dict(ntime=80, runtype="gradient", objective="Jmeasure", target="pure, 1, 0", initialcondition="diagonal",
                                                            penalty=0.5, penalty_param=1.5, penalty_energy=0.2, control_init="random, 0.004"), False),
    "C2_gradient_imr4_penalty": ("C2", dict(ntime=40, runtype="gradient", timestepper="IMR4", penalty=0.5, penalty_param=1.0, penalty_energy=0.1,
                                            control_init="random, 0.004"), False),
    "C3_simulation_imr8_outputs": ("C3", dict(ntime=12, runtype="simulation", timestepper="IMR8", control_init="random, 0.004",
                                               outputs="expectedEnergy, population", output_frequency=3), False),
    "C2_gradient_spline0_variation": ("C2", dict(ntime=60, runtype="gradient", segments="spline0, 12", penalty_variation=0.05, enforceBC=True,
                                                  control_init="random, 0.004"), False),
    "G33_gradient_two_segments": ("G33", dict(ntime=50, runtype="gradient", objective="Jfrobenius", target="pure, 1, 1", initialcondition="diagonal",
                                               segments="spline, 6, 0.0, 1.25, spline0, 5, 1.25, 2.5", penalty_variation=0.02,
                                               control_init="random, 0.004, random, 0.003"), False),
    "C1_gradient_dpdm_penalty": ("C1", dict(ntime=60, runtype="gradient", objective="Jmeasure", target="pure, 0, 1", initialcondition="diagonal",
                                            penalty_dpdm=0.02, penalty=0.3, penalty_param=1.0, control_init="random, 0.004"), False),
    "S2020_gradient_schroedinger_cnot_guard_levels": ("S2020", dict(ntime=40, runtype="gradient", objective="Jtrace", target="gate, cnot", nessential=[2, 2],
                                                                     penalty=0.5, penalty_param=0.5, penalty_dpdm=0.02, control_init="random, 0.003",
                                                                     outputs="expectedEnergy, population", output_frequency=10), False),
    "S888_gradient_schroedinger_measure_imr4": ("S888", dict(ntime=24, runtype="gradient", objective="Jmeasure", target="pure, 1, 0, 1", nessential=[2, 2, 2],
                                                               initialcondition="diagonal", timestepper="IMR4", penalty_energy=0.1, control_init="random, 0.003",
                                                               outputs="expectedEnergy", output_frequency=8, usematfree=False), False),
    "C1_gradient_penalty_jtrace_schroedinger": ("C1", dict(ntime=80, runtype="gradient", objective="Jtrace", target="gate, cnot", penalty=0.5, penalty_param=1.5,
                                                           control_init="random, 0.004"), False),
    "C1_gradient_penalty_jtrace_pure_imr4": ("C1", dict(ntime=30, runtype="gradient", objective="Jtrace", target="pure, 1, 0", initialcondition="diagonal",
                                                        timestepper="IMR4", penalty=0.8, penalty_param=1.0, control_init="random, 0.004"), False),
    "G33_gradient_decay_only": ("G33", dict(ntime=40, runtype="gradient", objective="Jfrobenius", target="pure, 2, 1", initialcondition="diagonal",
                                            control_init="random, 0.004", overrides={"collapse_type": "decay"}), False),
    "G33_gradient_dephase_only_nplus1": ("G33", dict(ntime=40, runtype="gradient", objective="Jmeasure", target="pure, 0, 1", initialcondition="Nplus1",
                                                      control_init="random, 0.004", overrides={"collapse_type": "dephase"},
                                                      outputs="population", output_frequency=10), False),
    "C1_gradient_gate_rotation_tik0": ("C1", dict(ntime=80, runtype="gradient", control_init="random, 0.004",
                                                  overrides={"gate_rot_freq": "0.05, -0.03", "optim_regul_tik0": "true", "optim_regul": "0.01"}), False),
    "C2_gradient_gate_rotation_swap": ("C2", dict(ntime=60, runtype="gradient", target="gate, swap", control_init="random, 0.004",
                                                   overrides={"gate_rot_freq": "0.02, 0.07"}), False),
    "C3_gradient_ensemble_subsystem": ("C3", dict(ntime=40, runtype="gradient", objective="Jmeasure", target="pure, 1, 0, 1", initialcondition="ensemble, 0, 1",
                                                   control_init="random, 0.004", outputs="expectedEnergy", output_frequency=10), False),
    "C3_gradient_basis_subsystem": ("C3", dict(ntime=40, runtype="gradient", objective="Jfrobenius", target="pure, 0, 1, 1", initialcondition="basis, 1",
                                                control_init="random, 0.004"), False),
    "C1_optimization_row0": ("C1", dict(ntime=100, runtype="optimization", control_init="constant, 0.005", optim_maxiter=0), False),
    # explicit Euler stepper (src/timestepper.cpp:484-519), Lindblad (stored trajectory) and Schroedinger (primal stepped backwards)
    "C2_gradient_explicit_euler": ("C2", dict(ntime=120, runtype="gradient", timestepper="EE", control_init="random, 0.004", penalty=0.3, penalty_param=2.0), False),
    "C1_gradient_explicit_euler_schroedinger": ("C1", dict(ntime=120, runtype="gradient", timestepper="EE", control_init="random, 0.004",
                                                            outputs="population", output_frequency=30), False),
    "C1_simulation_performance_initial_condition": ("C1", dict(ntime=80, runtype="simulation", initialcondition="performance", objective="Jmeasure",
                                                               target="pure, 1, 0", control_init="random, 0.004", outputs="population", output_frequency=20), False),
    # the configuration bench.py times (BASELINE config 4 at full length: 10 000 steps, GMRES, all 256 basis states, the
    # design vector bench.py's first evaluation uses); ~2 minutes on 8 cores; control files dropped (1.5 MB, pinned by the short cases)
    "C4_gradient_gmres_full": ("C4", dict(ntime=10000, runtype="gradient", linsolve="gmres"), True, dict(amp=0.005, np=8, drop=("control",))),
    # BASELINE config 5 through the entry point: three 4-level qudits with Jkl couplings, the N diagonal initial states,
    # pure target, Jmeasure
    "C5_gradient_measure_diagonal": ("C5", dict(ntime=200, runtype="gradient", objective="Jmeasure", target="pure, 0, 0, 0", initialcondition="diagonal"),
                                     True, dict(amp=0.005, np=8)),
}


def main():
    only = sys.argv[1:]          # case names: (re)generate just these; none: all
    if not only and os.path.exists(OUT):
        shutil.rmtree(OUT)
    for name, spec in CASES.items():
        system, kw, with_params = spec[:3]
        opt = spec[3] if len(spec) > 3 else {}
        if only and name not in only:
            continue
        d = os.path.join(OUT, name)
        if os.path.exists(d):
            shutil.rmtree(d)
        os.makedirs(d)
        if with_params:
            desc, _ = cfg_system(system, ntime=kw.get("ntime"))
            alpha = random_alpha(desc_ndesign(desc), amp=opt.get("amp", 0.004))
            with open(os.path.join(d, "params.dat"), "w") as f:
                f.write("\n".join(repr(float(v)) for v in alpha) + "\n")
        with open(os.path.join(d, "case.cfg"), "w") as f:
            f.write(cfg_text(system, **kw))
        env = dict(os.environ, QSHIM_NP=str(opt.get("np", 1)))
        r = subprocess.run([REF, "case.cfg", "--quiet"], cwd=d, env=env, capture_output=True, text=True)
        assert r.returncode == 0, (name, r.stderr[-400:])
        base = os.path.join(d, "base")
        os.rename(os.path.join(d, "data_out"), base)
        for f in ("config_log.dat", "timing.dat"):
            if os.path.exists(os.path.join(base, f)):
                os.remove(os.path.join(base, f))
        for f in os.listdir(base):
            if f.startswith(tuple(opt.get("drop", ()))):
                os.remove(os.path.join(base, f))
        print(name, sorted(os.listdir(base))[:6], "...")


if __name__ == "__main__":
    main()
