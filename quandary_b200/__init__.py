"""quandary_b200 -- thin ctypes binding of the C ABI in include/quandary_b200.h.

The product is the CUDA library `libquandary_b200.so` (built in-tree by `make -C quandary_b200` or
`__graft_entry__.build()`).  This module only loads it and mirrors the ABI one to one; there is no Python or CPU
fallback: importing works without a GPU (so that the ABI can be inspected), every compute call needs one.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libquandary_b200.so")

QB_MAX_OSC = 5
QB_MAX_PAIRS = 10
LINDBLAD = {"none": 0, "decay": 1, "dephase": 2, "both": 3}
LINSOLVE = {"gmres": 0, "neumann": 1}
CTRL = {"spline": 0, "spline0": 1}
STEPPER = {"IMR": 0, "IMR4": 4, "IMR8": 8}
FLAG_RECOMPUTE_STAGE = 1
FLAG_STORE_STAGE = 2
FLAG_ARNOLDI = 4
STORE_NONE, STORE_STATES, STORE_ADJOINT_INPUT = 0, 1, 2     # store_trajectory levels of qb_engine_forward[_dev]


class Segment(C.Structure):
    _fields_ = [("type", C.c_int32), ("nsplines", C.c_int32), ("tstart", C.c_double), ("tstop", C.c_double),
                ("skip", C.c_int32), ("enforce_bc", C.c_int32)]


class SystemDesc(C.Structure):
    _fields_ = [
        ("nosc", C.c_int32), ("nlevels", C.c_int32 * QB_MAX_OSC), ("lindblad", C.c_int32),
        ("detuning", C.c_double * QB_MAX_OSC), ("selfkerr", C.c_double * QB_MAX_OSC),
        ("crosskerr", C.c_double * QB_MAX_PAIRS), ("Jkl", C.c_double * QB_MAX_PAIRS), ("eta", C.c_double * QB_MAX_PAIRS),
        ("decay_time", C.c_double * QB_MAX_OSC), ("dephase_time", C.c_double * QB_MAX_OSC),
        ("ntime", C.c_int32), ("total_time", C.c_double), ("stepper", C.c_int32),
        ("linsolve_type", C.c_int32), ("linsolve_maxiter", C.c_int32),
        ("linsolve_abstol", C.c_double), ("linsolve_reltol", C.c_double),
        ("seg_off", C.c_int32 * (QB_MAX_OSC + 1)), ("segments", C.POINTER(Segment)),
        ("carrier_off", C.c_int32 * (QB_MAX_OSC + 1)), ("carrier_freq", C.POINTER(C.c_double)),
        ("pipulse_off", C.c_int32 * (QB_MAX_OSC + 1)), ("pipulse", C.POINTER(C.c_double)),
        ("device", C.c_int32), ("flags", C.c_int32),
    ]


class PenaltyDesc(C.Structure):
    _fields_ = [("a", C.POINTER(C.c_double)), ("c0", C.POINTER(C.c_double)), ("d", C.POINTER(C.c_double)), ("guard", C.POINTER(C.c_double)),
                ("wdt", C.POINTER(C.c_double)), ("jbar", C.POINTER(C.c_double)), ("ninit", C.c_int32), ("overlap", C.c_int32),
                ("purity", C.POINTER(C.c_double))]


class ObjectiveDesc(C.Structure):
    _fields_ = [("objective", C.c_int32), ("pure_id", C.c_int32), ("target", C.POINTER(C.c_double)), ("purity", C.POINTER(C.c_double)),
                ("weight", C.POINTER(C.c_double)), ("fid_scale", C.c_double), ("ninit", C.c_int32)]


OBJ_JFROBENIUS, OBJ_JTRACE, OBJ_JMEASURE = 0, 1, 2


class Stats(C.Structure):
    _fields_ = [("steps_fwd", C.c_int64), ("steps_adj", C.c_int64), ("rhs_evals", C.c_int64), ("linsolve_iters", C.c_int64),
                ("kernel_launches", C.c_int64), ("last_fwd_ms", C.c_double), ("last_adj_ms", C.c_double),
                ("total_fwd_ms", C.c_double), ("total_adj_ms", C.c_double), ("n_fwd", C.c_int64), ("n_adj", C.c_int64)]


def make_desc(nlevels, lindblad="none", detuning=None, selfkerr=None, crosskerr=None, Jkl=None, eta=None,
              decay_time=None, dephase_time=None, ntime=100, total_time=1.0, stepper="IMR", linsolve="gmres",
              maxiter=20, abstol=1e-10, reltol=1e-20, segments=None, carriers=None, pipulses=None, device=0, flags=0):
    """Builds a SystemDesc.  All frequencies ANGULAR (rad/ns).  segments[k] = list of dicts
    {type, nsplines, tstart, tstop, enforce_bc}; carriers[k] = list of angular carrier frequencies;
    pipulses[k] = list of (tstart, tstop, amp).  The returned object keeps the backing arrays alive."""
    Q = len(nlevels)
    npairs = Q * (Q - 1) // 2
    d = SystemDesc()
    d.nosc = Q
    for k in range(Q):
        d.nlevels[k] = int(nlevels[k])
        d.detuning[k] = float(detuning[k]) if detuning is not None else 0.0
        d.selfkerr[k] = float(selfkerr[k]) if selfkerr is not None else 0.0
        d.decay_time[k] = float(decay_time[k]) if decay_time is not None else 0.0
        d.dephase_time[k] = float(dephase_time[k]) if dephase_time is not None else 0.0
    for i in range(npairs):
        d.crosskerr[i] = float(crosskerr[i]) if crosskerr is not None else 0.0
        d.Jkl[i] = float(Jkl[i]) if Jkl is not None else 0.0
        d.eta[i] = float(eta[i]) if eta is not None else 0.0
    d.lindblad = LINDBLAD[lindblad] if isinstance(lindblad, str) else int(lindblad)
    d.ntime = int(ntime)
    d.total_time = float(total_time)
    d.stepper = STEPPER[stepper] if isinstance(stepper, str) else int(stepper)
    d.linsolve_type = LINSOLVE[linsolve] if isinstance(linsolve, str) else int(linsolve)
    d.linsolve_maxiter = int(maxiter)
    d.linsolve_abstol = float(abstol)
    d.linsolve_reltol = float(reltol)
    segments = segments if segments is not None else [[] for _ in range(Q)]
    carriers = carriers if carriers is not None else [[0.0] for _ in range(Q)]
    pipulses = pipulses if pipulses is not None else [[] for _ in range(Q)]
    segs, cars, pps = [], [], []
    for k in range(Q):
        d.seg_off[k] = len(segs)
        d.carrier_off[k] = len(cars)
        d.pipulse_off[k] = len(pps)
        skip = 0
        for s in segments[k]:
            sg = Segment()
            sg.type = CTRL[s.get("type", "spline")] if isinstance(s.get("type", "spline"), str) else int(s["type"])
            sg.nsplines = int(s["nsplines"])
            sg.tstart = float(s.get("tstart", 0.0))
            sg.tstop = float(s.get("tstop", total_time))
            sg.skip = skip
            sg.enforce_bc = int(bool(s.get("enforce_bc", False)))
            skip += 2 * sg.nsplines * len(carriers[k])
            segs.append(sg)
        cars.extend(float(c) for c in carriers[k])
        pps.extend(tuple(float(v) for v in p) for p in pipulses[k])
    d.seg_off[Q] = len(segs)
    d.carrier_off[Q] = len(cars)
    d.pipulse_off[Q] = len(pps)
    seg_arr = (Segment * max(1, len(segs)))(*segs)
    car_arr = (C.c_double * max(1, len(cars)))(*cars)
    flat_pp = [v for p in pps for v in p]
    pp_arr = (C.c_double * max(1, len(flat_pp)))(*flat_pp)
    d.segments = C.cast(seg_arr, C.POINTER(Segment))
    d.carrier_freq = C.cast(car_arr, C.POINTER(C.c_double))
    d.pipulse = C.cast(pp_arr, C.POINTER(C.c_double))
    d.device = int(device)
    d.flags = int(flags)
    d._keep = (seg_arr, car_arr, pp_arr)
    return d


def desc_dim(d):
    N = 1
    for k in range(d.nosc):
        N *= d.nlevels[k]
    return N if d.lindblad == 0 else N * N


def desc_ndesign(d):
    n = 0
    for k in range(d.nosc):
        nc = d.carrier_off[k + 1] - d.carrier_off[k]
        for s in range(d.seg_off[k], d.seg_off[k + 1]):
            n += 2 * d.segments[s].nsplines * nc
    return n


_lib = None


class QBError(RuntimeError):
    pass


def load_library(path=None):
    """Loads libquandary_b200.so.  Fails loudly if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise QBError(f"{p} not found: build it with `make -C {_HERE}` (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(p)
    dp = C.POINTER(C.c_double)
    vp = C.c_void_p
    sigs = {
        "qb_engine_create": ([C.POINTER(SystemDesc), C.POINTER(vp)], C.c_int),
        "qb_engine_destroy": ([vp], C.c_int),
        "qb_last_error": ([], C.c_char_p),
        "qb_engine_ndesign": ([vp], C.c_int),
        "qb_engine_dim": ([vp], C.c_int),
        "qb_engine_set_params": ([vp, dp, C.c_int], C.c_int),
        "qb_engine_eval_controls": ([vp, dp, C.c_int, dp, dp], C.c_int),
        "qb_engine_forward": ([vp, dp, C.c_int, C.c_int, dp], C.c_int),
        "qb_engine_adjoint": ([vp, dp, C.c_int, dp, dp], C.c_int),
        "qb_engine_get_state": ([vp, C.c_int, C.c_int, dp], C.c_int),
        "qb_engine_rhs_apply": ([vp, C.c_double, C.c_int, dp, dp, C.c_int], C.c_int),
        "qb_engine_upload_states": ([vp, dp, C.c_int, C.c_int], C.c_int),
        "qb_engine_forward_dev": ([vp, C.c_int, C.c_int], C.c_int),
        "qb_engine_adjoint_dev": ([vp, C.c_int], C.c_int),
        "qb_engine_download_states": ([vp, dp, C.c_int, C.c_int], C.c_int),
        "qb_engine_download_grad": ([vp, dp], C.c_int),
        "qb_engine_grad_dev": ([vp], vp),
        "qb_engine_set_stream": ([vp, vp, C.c_int], C.c_int),
        "qb_engine_set_penalty": ([vp, C.POINTER(PenaltyDesc)], C.c_int),
        "qb_engine_penalty_integrals": ([vp, C.c_int, dp], C.c_int),
        "qb_engine_set_objective": ([vp, C.POINTER(ObjectiveDesc)], C.c_int),
        "qb_engine_objective": ([vp, C.c_int, dp], C.c_int),
        "qb_engine_objective_bar": ([vp, C.c_int, C.c_double, C.c_double], C.c_int),
        "qb_engine_max_batch": ([vp, C.c_int], C.c_int),
        "qb_engine_set_dpdm": ([vp, dp, C.c_int], C.c_int),
        "qb_engine_dpdm_integrals": ([vp, C.c_int, dp], C.c_int),
        "qb_measure_fp64_peak": ([C.c_int, dp], C.c_int),
        "qb_measure_fp64_mma_peak": ([C.c_int, dp], C.c_int),
        "qb_engine_apply_bench": ([vp, C.c_double, C.c_int, C.c_int, C.c_int, dp], C.c_int),
        "qb_comm_unique_id": ([C.c_char_p], C.c_int),
        "qb_engine_comm_init": ([vp, C.c_char_p, C.c_int, C.c_int], C.c_int),
        "qb_engine_allreduce": ([vp, dp, C.c_int], C.c_int),
        "qb_engine_allreduce_grad_dev": ([vp], C.c_int),
        "qb_engine_sync": ([vp], C.c_int),
        "qb_engine_get_stats": ([vp, C.POINTER(Stats)], C.c_int),
    }
    for name, (argt, rest) in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = argt
        fn.restype = rest
    lib._qb_symbols = sorted(sigs)
    if path is None:
        _lib = lib
    return lib


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Engine:
    """One engine per CUDA device: the device-side MasterEq + TimeStepper for all initial conditions."""

    def __init__(self, desc):
        self.lib = load_library()
        self.desc = desc
        self.h = C.c_void_p()
        self._check(self.lib.qb_engine_create(C.byref(desc), C.byref(self.h)))
        self.dim = self.lib.qb_engine_dim(self.h)
        self.ndesign = self.lib.qb_engine_ndesign(self.h)

    def _check(self, rc):
        if rc != 0:
            raise QBError(f"quandary_b200 error {rc}: {self.lib.qb_last_error().decode()}")

    def close(self):
        if self.h:
            self.lib.qb_engine_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, alpha):
        a = np.ascontiguousarray(alpha, dtype=np.float64)
        self._check(self.lib.qb_engine_set_params(self.h, _dptr(a), a.size))

    def eval_controls(self, t):
        t = np.ascontiguousarray(t, dtype=np.float64)
        p = np.zeros((t.size, self.desc.nosc))
        q = np.zeros((t.size, self.desc.nosc))
        self._check(self.lib.qb_engine_eval_controls(self.h, _dptr(t), t.size, _dptr(p), _dptr(q)))
        return p, q

    def forward(self, rho0, store=False):
        r = np.ascontiguousarray(rho0, dtype=np.float64).reshape(-1, 2 * self.dim)
        out = np.empty_like(r)
        self._check(self.lib.qb_engine_forward(self.h, _dptr(r), r.shape[0], int(store), _dptr(out)))
        return out

    def adjoint(self, rhoT_bar, want_rho0_bar=False):
        r = np.ascontiguousarray(rhoT_bar, dtype=np.float64).reshape(-1, 2 * self.dim)
        g = np.zeros(max(1, self.ndesign))
        xb = np.empty_like(r) if want_rho0_bar else None
        self._check(self.lib.qb_engine_adjoint(self.h, _dptr(r), r.shape[0], _dptr(g), _dptr(xb) if want_rho0_bar else None))
        g = g[: self.ndesign]
        return (g, xb) if want_rho0_bar else g

    def get_state(self, iinit, tindex):
        out = np.empty(2 * self.dim)
        self._check(self.lib.qb_engine_get_state(self.h, iinit, tindex, _dptr(out)))
        return out

    def rhs_apply(self, t, x, transpose=False):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 2 * self.dim)
        y = np.empty_like(x)
        self._check(self.lib.qb_engine_rhs_apply(self.h, float(t), int(transpose), _dptr(x), _dptr(y), x.shape[0]))
        return y

    # device-resident variants
    def upload_states(self, states, which=0):
        r = np.ascontiguousarray(states, dtype=np.float64).reshape(-1, 2 * self.dim)
        self._check(self.lib.qb_engine_upload_states(self.h, _dptr(r), r.shape[0], which))
        return r.shape[0]

    def forward_dev(self, ninit, store=False):
        self._check(self.lib.qb_engine_forward_dev(self.h, ninit, int(store)))

    def adjoint_dev(self, ninit):
        self._check(self.lib.qb_engine_adjoint_dev(self.h, ninit))

    def download_states(self, ninit, which=0):
        out = np.empty((ninit, 2 * self.dim))
        self._check(self.lib.qb_engine_download_states(self.h, _dptr(out), ninit, which))
        return out

    def download_grad(self):
        g = np.zeros(max(1, self.ndesign))
        self._check(self.lib.qb_engine_download_grad(self.h, _dptr(g)))
        return g[: self.ndesign]

    def grad_dev_ptr(self):
        return self.lib.qb_engine_grad_dev(self.h)

    def set_penalty(self, a=None, c0=None, d=None, guard=None, wdt=None, jbar=None, overlap=False, purity=None):
        """installs the penalty functional (see include/quandary_b200.h); set_penalty() with no arguments removes it.
        overlap=True: Schroedinger + Jtrace form 1 - (Re<a,x>/purity)^2 - (Im<a,x>)^2 (c0, d unused)."""
        if a is None:
            self._check(self.lib.qb_engine_set_penalty(self.h, None))
            return
        arrs = [np.ascontiguousarray(v, dtype=np.float64) if v is not None else None for v in (a, c0, d, guard, wdt, jbar, purity)]
        pd = PenaltyDesc()
        pd.a, pd.wdt, pd.jbar = (_dptr(arrs[0]), _dptr(arrs[4]), _dptr(arrs[5]))
        pd.c0 = _dptr(arrs[1]) if arrs[1] is not None else None
        pd.d = _dptr(arrs[2]) if arrs[2] is not None else None
        pd.guard = _dptr(arrs[3]) if arrs[3] is not None else None
        pd.purity = _dptr(arrs[6]) if arrs[6] is not None else None
        pd.overlap = 1 if overlap else 0
        pd.ninit = arrs[5].size
        self._check(self.lib.qb_engine_set_penalty(self.h, C.byref(pd)))

    def penalty_integrals(self, ninit):
        out = np.zeros(ninit)
        self._check(self.lib.qb_engine_penalty_integrals(self.h, ninit, _dptr(out)))
        return out

    def set_objective(self, objective, weight, fid_scale, pure_id=-1, target=None, purity=None):
        """Final-time objective on the device (OptimTarget::evalJ & co., src/optimtarget.cpp:343-447, 713-862); None removes it."""
        if objective is None:
            self._check(self.lib.qb_engine_set_objective(self.h, None))
            return
        w = np.ascontiguousarray(weight, dtype=np.float64)
        od = ObjectiveDesc()
        od.objective, od.pure_id, od.fid_scale, od.ninit = int(objective), int(pure_id), float(fid_scale), w.shape[0]
        od.weight = _dptr(w)
        keep = [w]
        if target is not None:
            t = np.ascontiguousarray(target, dtype=np.float64).reshape(w.shape[0], 2 * self.dim)
            od.target = _dptr(t)
            keep.append(t)
        if purity is not None:
            pu = np.ascontiguousarray(purity, dtype=np.float64).reshape(w.shape[0])
            od.purity = _dptr(pu)
            keep.append(pu)
        self._check(self.lib.qb_engine_set_objective(self.h, C.byref(od)))

    def objective(self, ninit):
        """(sum w Re J, sum w Im J, fid_scale sum Re HS, fid_scale sum Im HS) of the last forward sweep's final states."""
        out = np.zeros(4)
        self._check(self.lib.qb_engine_objective(self.h, ninit, _dptr(out)))
        return out

    def objective_bar(self, ninit, re_bar, im_bar=0.0):
        """Writes the terminal adjoint states (evalJ_diff) into the device buffer qb_engine_adjoint_dev starts from."""
        self._check(self.lib.qb_engine_objective_bar(self.h, ninit, float(re_bar), float(im_bar)))

    def comm_init(self, nccl_id, rank, nranks):
        """joins the NCCL communicator of the initial-condition sharding group (qb_engine_comm_init); nccl_id: 128 bytes from comm_unique_id() of rank 0"""
        self._check(self.lib.qb_engine_comm_init(self.h, bytes(nccl_id), rank, nranks))

    def allreduce(self, buf):
        """in-place SUM of a host array of at most ndesign + 8 doubles over the ranks (qb_engine_allreduce)"""
        b = np.ascontiguousarray(buf, dtype=np.float64)
        self._check(self.lib.qb_engine_allreduce(self.h, _dptr(b), b.size))
        return b

    def allreduce_grad_dev(self):
        """in-place NCCL all-reduce of the device gradient buffer on the engine's stream (qb_engine_allreduce_grad_dev)"""
        self._check(self.lib.qb_engine_allreduce_grad_dev(self.h))

    def apply_bench(self, t, nbatch, reps, dense):
        """device time (ms) of one launch applying M(t) `reps` times to `nbatch` resident states; dense: the DMMA formulation (N = 64)"""
        ms = C.c_double()
        self._check(self.lib.qb_engine_apply_bench(self.h, float(t), int(nbatch), int(reps), int(bool(dense)), C.byref(ms)))
        return ms.value

    def set_stream(self, cuda_stream=None):
        """cuda_stream: integer cudaStream_t handle (e.g. torch.cuda.current_stream().cuda_stream); None = engine-owned."""
        if cuda_stream is None:
            self._check(self.lib.qb_engine_set_stream(self.h, None, 0))
        else:
            self._check(self.lib.qb_engine_set_stream(self.h, C.c_void_p(int(cuda_stream)), 1))

    def sync(self):
        self._check(self.lib.qb_engine_sync(self.h))

    def stats(self):
        s = Stats()
        self._check(self.lib.qb_engine_get_stats(self.h, C.byref(s)))
        return {f[0]: getattr(s, f[0]) for f in Stats._fields_}


def comm_unique_id():
    """128-byte NCCL id created by rank 0 (qb_comm_unique_id); hand it to the other ranks over any host channel"""
    lib = load_library()
    buf = C.create_string_buffer(128)
    if lib.qb_comm_unique_id(buf) != 0:
        raise QBError(lib.qb_last_error().decode())
    return buf.raw


def measure_fp64_mma_peak(device=0):
    """fp64 tensor-pipe (DMMA, mma.sync m8n8k4) peak of the device in TFLOP/s"""
    lib = load_library()
    v = C.c_double()
    if lib.qb_measure_fp64_mma_peak(device, C.byref(v)) != 0:
        raise QBError(lib.qb_last_error().decode())
    return v.value


def measure_fp64_peak(device=0):
    """Measured DFMA peak (TFLOP/s) of the device: the FP64 roofline denominator."""
    lib = load_library()
    v = C.c_double(0.0)
    rc = lib.qb_measure_fp64_peak(int(device), C.byref(v))
    if rc != 0:
        raise QBError(f"quandary_b200 error {rc}: {lib.qb_last_error().decode()}")
    return v.value
