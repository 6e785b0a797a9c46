"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

ctypes front-end over the two CPU oracles (see oracle_api.h):

* ``impl="port"``  -> oracle/libcyrk_oracle.so  (plain-C restatement, always available)
* ``impl="ref"``   -> oracle/_ref/libcyrk_ref.so (the unmodified reference compiled from
  /root/reference by oracle/Makefile; prebuilt file travels to the GPU box)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference``
legs may import this module.  The product package (cyrk_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libcyrk_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libcyrk_ref.so")

METHODS = {"rk23": 3, "rk45": 4, "dop853": 5}  # cysolver.hpp:34-41

RHS = dict(baseline=0, accuracy=1, extraoutput=2, lorenz=3, lorenz_extraoutput=4,
           lotkavolterra=5, pendulum=6, arbitrary_args=7, pendulum_preeval=8,
           largey=9, largey_simple=10, pendulum_drag=11, dense_kat=12)
EVENTS = dict(y0_gt_50=0, y1_lt_50=1, t_gt_5=2, lorenz_t_ge_5=3, lorenz_y_range=4,
              lorenz_args=5, y0=6, y1=7, t_minus_arg4=8, abs_y0_minus_arg4=9)

SIZE_MAX = (1 << 64) - 1

DIFFEQ_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double), C.c_void_p, C.c_void_p)


class _Problem(C.Structure):
    _fields_ = [
        ("diffeq", C.c_void_p), ("pre_eval", C.c_void_p), ("method", C.c_int),
        ("t_start", C.c_double), ("t_end", C.c_double),
        ("y0", C.c_void_p), ("num_y", C.c_size_t), ("num_extra", C.c_size_t),
        ("args", C.c_void_p), ("args_len", C.c_size_t),
        ("max_num_steps", C.c_size_t), ("max_ram_MB", C.c_size_t), ("expected_size", C.c_size_t),
        ("dense_output", C.c_int),
        ("t_eval", C.c_void_p), ("n_t_eval", C.c_size_t),
        ("event_fns", C.c_void_p), ("event_dirs", C.c_void_p), ("event_max", C.c_void_p),
        ("n_events", C.c_size_t),
        ("rtols", C.c_void_p), ("n_rtols", C.c_size_t),
        ("atols", C.c_void_p), ("n_atols", C.c_size_t),
        ("max_step", C.c_double), ("first_step", C.c_double),
        ("fix_quirk11", C.c_int),
    ]


def build(ref: bool = True) -> None:
    """Compile the C restatement (always) and oracle/_ref (when /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", HERE, "port"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


_LIBS: dict = {}


def _load(impl: str):
    if impl in _LIBS:
        return _LIBS[impl]
    path, pfx = (PORT_SO, "orc") if impl == "port" else (REF_SO, "ref")
    if not os.path.exists(path):
        if impl == "port":
            build(ref=False)
        else:
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
    lib = C.CDLL(path)
    f = lambda n: getattr(lib, f"{pfx}_{n}")
    f("solve").restype = C.c_void_p
    f("solve").argtypes = [C.POINTER(_Problem)]
    f("free").argtypes = [C.c_void_p]
    for n in ("status", "success", "event_terminated"):
        f(n).restype = C.c_int
        f(n).argtypes = [C.c_void_p]
    for n in ("size", "steps_taken", "num_interpolates", "num_dy", "nfev", "event_terminate_index"):
        f(n).restype = C.c_size_t
        f(n).argtypes = [C.c_void_p]
    for n in ("t_ptr", "y_ptr"):
        f(n).restype = C.POINTER(C.c_double)
        f(n).argtypes = [C.c_void_p]
    f("event_count").restype = C.c_size_t
    f("event_count").argtypes = [C.c_void_p, C.c_size_t]
    for n in ("event_t_ptr", "event_y_ptr"):
        f(n).restype = C.POINTER(C.c_double)
        f(n).argtypes = [C.c_void_p, C.c_size_t]
    f("call").restype = C.c_int
    f("call").argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    f("solve_batch").restype = C.c_double
    f("solve_batch").argtypes = [C.POINTER(_Problem), C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                 C.c_void_p, C.c_int] + [C.c_void_p] * 7
    f("brentq_poly3").restype = C.c_double
    f("brentq_poly3").argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_size_t,
                                  C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.orc_rhs_get.restype = C.c_void_p
    lib.orc_rhs_get.argtypes = [C.c_int]
    lib.orc_rhs_pre_eval.restype = C.c_void_p
    lib.orc_rhs_pre_eval.argtypes = [C.c_int]
    lib.orc_rhs_num_extra.restype = C.c_int
    lib.orc_rhs_num_extra.argtypes = [C.c_int]
    lib.orc_event_get.restype = C.c_void_p
    lib.orc_event_get.argtypes = [C.c_int]
    _LIBS[impl] = (lib, f)
    return _LIBS[impl]


def have_ref() -> bool:
    return os.path.exists(REF_SO)


@dataclass
class OracleResult:
    status: int
    success: bool
    size: int
    steps_taken: int
    num_interpolates: int
    nfev: int
    t: np.ndarray          # (size,)
    y: np.ndarray          # (num_dy, size)  -- the reference's Python-view shape (cysolver_api.pyx:273-275)
    event_terminated: bool = False
    event_terminate_index: int = SIZE_MAX
    t_events: list = field(default_factory=list)
    y_events: list = field(default_factory=list)   # each (num_dy, hits)
    dense: Optional[np.ndarray] = None             # (num_dy, n_call) if call_t was given
    dense_code: int = 0

    @property
    def n_attempts(self):
        return None


def _as_args(args) -> bytes:
    if args is None:
        return b""
    if isinstance(args, (bytes, bytearray)):
        return bytes(args)
    return np.ascontiguousarray(args, dtype=np.float64).tobytes()


class _Prepared:
    """Keeps the numpy buffers alive for the duration of a call."""

    def __init__(self, lib, rhs, method, t_span, y0, args, rtol, atol, t_eval, dense_output, events,
                 num_extra, first_step, max_step, max_num_steps, max_ram_MB, expected_size, fix_quirk11):
        p = _Problem()
        if isinstance(rhs, str):
            rid = RHS[rhs]
            p.diffeq = lib.orc_rhs_get(rid)
            p.pre_eval = lib.orc_rhs_pre_eval(rid)
            if num_extra is None:
                num_extra = lib.orc_rhs_num_extra(rid)
        elif isinstance(rhs, int):
            p.diffeq = rhs            # raw function address (e.g. numba cfunc)
            p.pre_eval = None
        else:                          # ctypes callback
            self._cb = rhs
            p.diffeq = C.cast(rhs, C.c_void_p).value
            p.pre_eval = None
        p.method = METHODS[method.lower()] if isinstance(method, str) else int(method)
        p.t_start, p.t_end = float(t_span[0]), float(t_span[1])
        self.y0 = np.ascontiguousarray(y0, dtype=np.float64)
        p.y0 = self.y0.ctypes.data
        p.num_y = self.y0.shape[-1]
        p.num_extra = int(num_extra or 0)
        self.args = _as_args(args)
        self.args_buf = C.create_string_buffer(self.args, max(len(self.args), 1))
        p.args = C.addressof(self.args_buf)
        p.args_len = len(self.args)
        p.max_num_steps, p.max_ram_MB, p.expected_size = int(max_num_steps), int(max_ram_MB), int(expected_size)
        p.dense_output = int(bool(dense_output))
        self.t_eval = np.ascontiguousarray(t_eval if t_eval is not None else [], dtype=np.float64)
        p.t_eval = self.t_eval.ctypes.data
        p.n_t_eval = self.t_eval.size
        events = events or []
        self.ev_fns = (C.c_void_p * max(len(events), 1))()
        self.ev_dirs = np.zeros(max(len(events), 1), dtype=np.int32)
        self.ev_max = np.full(max(len(events), 1), SIZE_MAX, dtype=np.uint64)
        for i, ev in enumerate(events):
            name, direction, max_allowed = (ev + (0, None))[:3] if isinstance(ev, tuple) else (ev, 0, None)
            self.ev_fns[i] = lib.orc_event_get(EVENTS[name]) if isinstance(name, str) else name
            self.ev_dirs[i] = direction
            self.ev_max[i] = SIZE_MAX if max_allowed is None else max_allowed
        p.event_fns = C.addressof(self.ev_fns)
        p.event_dirs = self.ev_dirs.ctypes.data
        p.event_max = self.ev_max.ctypes.data
        p.n_events = len(events)
        self.rtols = np.atleast_1d(np.asarray(rtol, dtype=np.float64)).copy()
        self.atols = np.atleast_1d(np.asarray(atol, dtype=np.float64)).copy()
        p.rtols, p.n_rtols = self.rtols.ctypes.data, self.rtols.size
        p.atols, p.n_atols = self.atols.ctypes.data, self.atols.size
        p.max_step, p.first_step = float(max_step), float(first_step)
        p.fix_quirk11 = int(fix_quirk11)
        self.p = p
        self.num_dy = p.num_y + p.num_extra


def solve(rhs, t_span, y0, method="RK45", rtol=1e-3, atol=1e-6, args=None, t_eval=None,
          dense_output=False, events=None, num_extra=None, first_step=0.0, max_step=np.inf,
          max_num_steps=0, max_ram_MB=2000, expected_size=0, call_t=None, impl="port",
          fix_quirk11=False) -> OracleResult:
    """One solve with the argument surface of cysolve_ivp (cysolver_api.pyx:416-469).

    ``events``: list of names or (name, direction, max_allowed) tuples from ``EVENTS``.
    """
    lib, f = _load(impl)
    pr = _Prepared(lib, rhs, method, t_span, y0, args, rtol, atol, t_eval, dense_output, events,
                   num_extra, first_step, max_step, max_num_steps, max_ram_MB, expected_size, fix_quirk11)
    h = f("solve")(C.byref(pr.p))
    try:
        size = f("size")(h)
        ndy = f("num_dy")(h)
        t = np.ctypeslib.as_array(f("t_ptr")(h), shape=(max(size, 1),))[:size].copy()
        y = np.ctypeslib.as_array(f("y_ptr")(h), shape=(max(size * ndy, 1),))[: size * ndy].copy()
        res = OracleResult(
            status=f("status")(h), success=bool(f("success")(h)), size=size,
            steps_taken=f("steps_taken")(h), num_interpolates=f("num_interpolates")(h), nfev=f("nfev")(h),
            t=t, y=y.reshape(size, ndy).T.copy(),
            event_terminated=bool(f("event_terminated")(h)),
            event_terminate_index=f("event_terminate_index")(h))
        for e in range(pr.p.n_events):
            n = f("event_count")(h, e)
            if n:
                te = np.ctypeslib.as_array(f("event_t_ptr")(h, e), shape=(n,)).copy()
                ye = np.ctypeslib.as_array(f("event_y_ptr")(h, e), shape=(n * ndy,)).copy().reshape(n, ndy).T.copy()
            else:
                te, ye = np.empty(0), np.empty((ndy, 0))
            res.t_events.append(te)
            res.y_events.append(ye)
        if call_t is not None:
            ct = np.atleast_1d(np.asarray(call_t, dtype=np.float64)).copy()
            out = np.full((ct.size, ndy), np.nan)
            res.dense_code = f("call")(h, ct.ctypes.data, ct.size, out.ctypes.data)
            res.dense = out.T.copy()
        return res
    finally:
        f("free")(h)


def solve_batch(rhs, t_span, y0, method="RK45", rtol=1e-3, atol=1e-6, args=None, t_eval=None,
                dense_output=False, events=None, num_extra=None, first_step=0.0, max_step=np.inf,
                max_num_steps=0, max_ram_MB=2000, expected_size=0, impl="port", n_threads=None,
                fix_quirk11=False, want_teval=True):
    """Loop of independent solves (prange pattern).  y0: (B, n_y); args: (n_args,) or (B, n_args);
    t_span: (2,) or (B, 2).  Returns dict of per-trajectory arrays + 'seconds'."""
    lib, f = _load(impl)
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    B = y0.shape[0]
    a = None if args is None else np.ascontiguousarray(args, dtype=np.float64)
    per_traj_args = a is not None and a.ndim == 2
    ts = np.ascontiguousarray(t_span, dtype=np.float64)
    per_traj_span = ts.ndim == 2
    pr = _Prepared(lib, rhs, method, ts[0] if per_traj_span else ts, y0[0], a[0] if per_traj_args else a,
                   rtol, atol, t_eval, dense_output, events, num_extra, first_step, max_step,
                   max_num_steps, max_ram_MB, expected_size, fix_quirk11)
    ndy = pr.num_dy
    out = dict(status=np.zeros(B, np.int32), size=np.zeros(B, np.uint64), steps=np.zeros(B, np.uint64),
               nfev=np.zeros(B, np.uint64), t_last=np.full(B, np.nan), y_last=np.full((B, ndy), np.nan))
    te = None
    if want_teval and pr.t_eval.size:
        te = np.full((B, pr.t_eval.size, ndy), np.nan)
    if n_threads is None:
        n_threads = len(os.sched_getaffinity(0))
    sec = f("solve_batch")(
        C.byref(pr.p), B, y0.ctypes.data,
        a.ctypes.data if per_traj_args else None, a.shape[1] * 8 if per_traj_args else 0,
        ts.ctypes.data if per_traj_span else None, int(n_threads),
        out["status"].ctypes.data, out["size"].ctypes.data, out["steps"].ctypes.data, out["nfev"].ctypes.data,
        out["t_last"].ctypes.data, out["y_last"].ctypes.data, te.ctypes.data if te is not None else None)
    out["seconds"] = sec
    out["n_threads"] = int(n_threads)
    if te is not None:
        out["teval_y"] = te
    return out


def brentq_poly3(a, xa, xb, xtol, rtol, mitr, impl="port"):
    lib, f = _load(impl)
    a = np.ascontiguousarray(a, dtype=np.float64)
    err, fc, it = C.c_int(0), C.c_size_t(0), C.c_size_t(0)
    r = f("brentq_poly3")(a.ctypes.data, xa, xb, xtol, rtol, mitr, C.byref(err), C.byref(fc), C.byref(it))
    return r, err.value, fc.value, it.value
