"""ctypes loader of oracle/libqoracle.so -- the CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE:
used by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg only, never by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

from quandary_b200 import SystemDesc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "libqoracle.so")

_dp = C.POINTER(C.c_double)
_lp = C.POINTER(C.c_long)


def _p(a):
    return a.ctypes.data_as(_dp)


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        dref = C.POINTER(SystemDesc)
        lib.qo_dim.argtypes = [dref]; lib.qo_dim.restype = C.c_int
        lib.qo_ndesign.argtypes = [dref]; lib.qo_ndesign.restype = C.c_int
        lib.qo_eval_control.argtypes = [dref, _dp, C.c_int, C.c_double, _dp, _dp]; lib.qo_eval_control.restype = None
        lib.qo_eval_control_diff.argtypes = [dref, C.c_int, C.c_double, _dp, C.c_double, C.c_double]; lib.qo_eval_control_diff.restype = None
        lib.qo_rhs_apply.argtypes = [dref, _dp, C.c_double, C.c_int, _dp, _dp]; lib.qo_rhs_apply.restype = None
        lib.qo_drhs_dparams.argtypes = [dref, C.c_double, _dp, _dp, C.c_double, _dp]; lib.qo_drhs_dparams.restype = None
        lib.qo_solve_ode.argtypes = [dref, _dp, _dp, _dp, _dp, _lp]; lib.qo_solve_ode.restype = C.c_long
        lib.qo_solve_adjoint.argtypes = [dref, _dp, _dp, _dp, _dp, _dp, _dp, _lp]; lib.qo_solve_adjoint.restype = C.c_long
        lib.qo_batch_forward.argtypes = [dref, _dp, _dp, C.c_int, _dp, C.c_int, _lp]; lib.qo_batch_forward.restype = C.c_long
        lib.qo_batch_gradient.argtypes = [dref, _dp, _dp, C.c_int, _dp, _dp, C.c_void_p, C.c_void_p, _dp, C.c_int, _lp]
        lib.qo_batch_gradient.restype = C.c_long
        lib.qo_num_threads.argtypes = []; lib.qo_num_threads.restype = C.c_int

    def dim(self, d):
        return self.lib.qo_dim(C.byref(d))

    def ndesign(self, d):
        return self.lib.qo_ndesign(C.byref(d))

    def eval_control(self, d, alpha, osc, t):
        a = np.ascontiguousarray(alpha, dtype=np.float64)
        p, q = C.c_double(), C.c_double()
        self.lib.qo_eval_control(C.byref(d), _p(a), osc, t, C.byref(p), C.byref(q))
        return p.value, q.value

    def rhs_apply(self, d, alpha, t, x, transpose=False):
        a = np.ascontiguousarray(alpha, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 2 * self.dim(d))
        y = np.empty_like(x)
        for i in range(x.shape[0]):
            self.lib.qo_rhs_apply(C.byref(d), _p(a), t, int(transpose), _p(x[i]), _p(y[i]))
        return y

    def drhs_dparams(self, d, t, x, xbar, scale=1.0):
        g = np.zeros(max(1, self.ndesign(d)))
        x = np.ascontiguousarray(x, dtype=np.float64); xbar = np.ascontiguousarray(xbar, dtype=np.float64)
        self.lib.qo_drhs_dparams(C.byref(d), t, _p(x), _p(xbar), scale, _p(g))
        return g[: self.ndesign(d)]

    def solve_ode(self, d, alpha, rho0, want_traj=False):
        a = np.ascontiguousarray(alpha, dtype=np.float64)
        n2 = 2 * self.dim(d)
        r0 = np.ascontiguousarray(rho0, dtype=np.float64).reshape(n2)
        rT = np.empty(n2)
        traj = np.empty((d.ntime + 1, n2)) if want_traj else None
        ev = C.c_long(0)
        it = self.lib.qo_solve_ode(C.byref(d), _p(a), _p(r0), _p(rT), _p(traj) if want_traj else None, C.byref(ev))
        return rT, traj, it, ev.value

    def solve_adjoint(self, d, alpha, rhoT_bar, finalstate, traj):
        a = np.ascontiguousarray(alpha, dtype=np.float64)
        n2 = 2 * self.dim(d)
        g = np.zeros(max(1, self.ndesign(d)))
        xb0 = np.empty(n2)
        ev = C.c_long(0)
        rb = np.ascontiguousarray(rhoT_bar, dtype=np.float64).reshape(n2)
        fs = np.ascontiguousarray(finalstate, dtype=np.float64).reshape(n2).copy()
        it = self.lib.qo_solve_adjoint(C.byref(d), _p(a), _p(rb), _p(fs), _p(traj) if traj is not None else None, _p(g), _p(xb0), C.byref(ev))
        return g[: self.ndesign(d)], xb0, it, ev.value

    def batch_forward(self, d, alpha, rho0, nthreads=0):
        a = np.ascontiguousarray(alpha, dtype=np.float64)
        r0 = np.ascontiguousarray(rho0, dtype=np.float64).reshape(-1, 2 * self.dim(d))
        rT = np.empty_like(r0)
        ev = C.c_long(0)
        it = self.lib.qo_batch_forward(C.byref(d), _p(a), _p(r0), r0.shape[0], _p(rT), nthreads, C.byref(ev))
        return rT, it, ev.value

    def batch_gradient(self, d, alpha, rho0, rhoT_bar, nthreads=0):
        """forward + adjoint for every initial condition with given terminal adjoints; returns (rhoT, grad, iters, evals)"""
        a = np.ascontiguousarray(alpha, dtype=np.float64)
        r0 = np.ascontiguousarray(rho0, dtype=np.float64).reshape(-1, 2 * self.dim(d))
        rb = np.ascontiguousarray(rhoT_bar, dtype=np.float64).reshape(-1, 2 * self.dim(d))
        rT = np.empty_like(r0)
        g = np.zeros(max(1, self.ndesign(d)))
        ev = C.c_long(0)
        it = self.lib.qo_batch_gradient(C.byref(d), _p(a), _p(r0), r0.shape[0], _p(rT), _p(rb), None, None, _p(g), nthreads, C.byref(ev))
        return rT, g[: self.ndesign(d)], it, ev.value

    def num_threads(self):
        return self.lib.qo_num_threads()


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "port"], check=True, capture_output=True)


def load():
    if not os.path.exists(LIB):
        build()
    return Oracle(C.CDLL(LIB))
