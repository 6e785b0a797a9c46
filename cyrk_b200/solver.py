"""Host-side mirror of the reference call surface for the batched GPU path.

``cysolve_ivp_batch(diffeq, t_span, y0, ...)`` keeps the argument names, defaults and error
behaviour of ``pysolve_ivp`` / ``cysolve_ivp`` (CyRK/cy/pysolver.pyx:348-403,
CyRK/cy/cysolver_api.pyx:416-469) but takes a *batch* of initial conditions ``y0[B, n_y]``
(and optionally per-trajectory ``args[B, n_args]`` / ``t_span[B, 2]``) and runs them in one
launch through the C ABI (include/cyrk_b200.h).  ``result[i]`` is a view with the attributes of
the reference's ``WrapCySolverResult`` (cysolver_api.pyx:244-357).

A Python callable cannot run on the GPU, so ``diffeq`` is the *name* of a built-in device functor
(the reference's benchmark/test problems) or a :class:`UserRHS` compiled from CUDA source.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Union

import numpy as np

from . import _lib
from ._lib import CyrkB200Error

INF = float("inf")
SIZE_MAX = (1 << 64) - 1

# pyhelpers.find_ode_method_int (CyRK/cy/pyhelpers.py:13-23)
_METHODS = {"rk23": 3, "rk45": 4, "dop853": 5}
_METHOD_DESCR = {   # CyrkODEMethods, cysolver.hpp:48-66
    3: "Explicit Runge-Kutta method of order 3(2).",
    4: "Explicit Runge-Kutta method of order 5(4)",
    5: "Explicit Runge-Kutta method of order 8. Error is controlled using a combination of 5th and 3rd order interpolators.",
}

EVENT_IDS = dict(y0_gt_50=0, y1_lt_50=1, t_gt_5=2, lorenz_t_ge_5=3, lorenz_y_range=4, lorenz_args=5,
                 y0=6, y1=7, t_minus_arg4=8, abs_y0_minus_arg4=9)

STORE_FINAL, STORE_ROWS = 0, 1
FLAG_STRICT_FP, FLAG_DEVICE_IO, FLAG_T_EVAL_PER_TRAJ, FLAG_TOL_PER_TRAJ = 1, 2, 4, 8


def find_ode_method_int(method: str) -> int:
    m = _METHODS.get(str(method).lower())
    if m is None:
        # same exception type/message family as pysolver.pyx:83-87
        raise NotImplementedError(
            f"ERROR: `cysolve_ivp_batch` - Unknown or unsupported integration method provided: {method}.\n"
            f"Supported methods are: RK23, RK45, DOP853.")
    return m


class Event:
    """Event description: mirrors the attributes the reference reads off event callables
    (``terminal``, ``max_allowed``, ``direction``; pysolver.pyx:152-200).

    ``func`` is the name of a built-in event functor (see ``EVENT_IDS``) or, for a
    :class:`UserRHS`, the integer index of an event compiled with it.
    """

    def __init__(self, func: Union[str, int], direction: int = 0, terminal: bool = False,
                 max_allowed: Optional[int] = None):
        self.func = func
        self.direction = int(direction)
        if max_allowed is None:
            max_allowed = 1 if terminal else None
        self.max_allowed = max_allowed
        self.terminal = bool(terminal) or (max_allowed is not None)


class UserRHS:
    """A user right-hand side compiled from CUDA C++ with NVRTC (cyrk_b200_rhs_compile).

    The source defines ``__device__ void <entry>(double* dy, double t, const double* y, const char* args)``
    -- the reference's DiffeqFuncType without the pre-eval pointer (c_common.hpp:218-220) -- and
    optionally ``__device__ double <entry>_event(int id, double t, const double* y, const char* args)``.
    """

    def __init__(self, source: str, entry: str, n_y: int, n_extra: int = 0, args_bytes: int = 0, n_events: int = 0):
        log = C.create_string_buffer(1 << 16)
        rid = _lib.lib().cyrk_b200_rhs_compile(source.encode(), entry.encode(), n_y, n_extra, args_bytes, n_events,
                                                log, len(log))
        self.log = log.value.decode("utf-8", "replace")
        if rid < 0:
            raise CyrkB200Error(f"NVRTC compilation failed: {_lib.last_error()}\n{self.log}")
        self.rhs_id = rid
        self.n_y, self.n_extra, self.args_bytes, self.n_events = n_y, n_extra, args_bytes, n_events


def _rhs_id(diffeq) -> int:
    if isinstance(diffeq, UserRHS):
        return diffeq.rhs_id
    if isinstance(diffeq, (int, np.integer)):
        return int(diffeq)
    if isinstance(diffeq, str):
        rid = _lib.lib().cyrk_b200_rhs_builtin(diffeq.encode())
        if rid < 0:
            raise NotImplementedError(f"unknown built-in right-hand side {diffeq!r}")
        return rid
    raise TypeError(
        "diffeq must be the name of a built-in device functor or a UserRHS: a Python callable cannot run on the GPU")


class DeviceArray:
    """A device-resident result array (CYRK_FLAG_DEVICE_IO).  Implements ``__cuda_array_interface__`` so
    ``torch.as_tensor(arr, device='cuda')`` / cupy wrap it without a copy.  Valid until the owning
    BatchResult is closed or re-used."""

    def __init__(self, ptr, shape, dtype, owner):
        self.ptr, self.shape, self.dtype, self._owner = int(ptr or 0), tuple(int(x) for x in shape), np.dtype(dtype), owner

    @property
    def __cuda_array_interface__(self):
        return dict(shape=self.shape, typestr=self.dtype.str, data=(self.ptr, False), version=3, strides=None)


def _is_cuda_tensor(x) -> bool:
    return hasattr(x, "is_cuda") and bool(getattr(x, "is_cuda")) and hasattr(x, "data_ptr")


class _HandleBox:
    """Owns one native result handle (cyrk_b200_free on collection).  Every array view a BatchResult hands out keeps a
    reference to its box, so the pinned memory under a view outlives the BatchResult it came from; the box is freed
    when the last result AND the last view are gone, or explicitly by ``close()``."""

    def __init__(self, h):
        self.h = h

    def close(self):
        if self.h:
            _lib.lib().cyrk_b200_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _OwnedArray(np.ndarray):
    """Zero-copy view onto pinned memory of a native handle; `_owner` (propagated to slices and reshapes) keeps the
    handle alive.  Valid until the owning result is closed or handed to ``solution_reuse`` -- copy (``np.array(x)``) what
    must survive that, exactly like the reference's reused CySolverResult (Documentation/CySolverResult_Reuses.md)."""
    _owner = None

    def __array_finalize__(self, obj):
        if obj is not None:
            self._owner = getattr(obj, "_owner", None)


def _np_view(ptr, shape, dtype, owner=None):
    n = int(np.prod(shape)) if len(shape) else 1
    if not ptr or n == 0:
        return np.empty(shape, dtype=dtype)
    ctype = {np.float64: C.c_double, np.int32: C.c_int32, np.uint32: C.c_uint32, np.uint64: C.c_uint64}[dtype]
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,)).reshape(shape).view(_OwnedArray)
    arr._owner = owner
    return arr


class TrajectoryView:
    """One trajectory of a :class:`BatchResult` with the WrapCySolverResult attribute surface."""

    def __init__(self, batch: "BatchResult", i: int):
        self._b, self._i = batch, int(i)

    # --- scalars
    @property
    def status(self): return int(self._b.status[self._i])
    error_code = status
    @property
    def success(self): return self.status in (1, 2)
    @property
    def message(self): return _lib.status_message(self.status)
    status_message = message
    @property
    def size(self): return int(self._b.size[self._i])
    @property
    def steps_taken(self): return int(self._b.steps_taken[self._i])
    @property
    def num_y(self): return self._b.num_y
    @property
    def num_dy(self): return self._b.num_dy
    @property
    def num_events(self): return self._b.num_events
    @property
    def integration_method(self): return _METHOD_DESCR[self._b.method]
    method = integration_method

    # --- arrays (copies, like the reference: cysolver_api.pyx:269-275)
    @property
    def t(self):
        b = self._b
        if b.t_rows is None:
            raise CyrkB200Error("rows were not stored: call cysolve_ivp_batch(..., store='rows')")
        return np.array(b.t_rows[self._i, :self.size])

    @property
    def y(self):
        b = self._b
        if b.y_rows is None:
            raise CyrkB200Error("rows were not stored: call cysolve_ivp_batch(..., store='rows')")
        return np.array(b.y_rows[self._i, :self.size, :]).T      # (num_dy, size)

    @property
    def t_events(self):
        b = self._b
        if b.num_events == 0:
            return None
        if b.events_packed:
            off = b.event_offsets
            return [np.array(b.event_t[off[self._i * b.num_events + e]:off[self._i * b.num_events + e + 1]])
                    for e in range(b.num_events)]
        return [np.array(b.event_t[self._i, e, :b.event_count[self._i, e]]) for e in range(b.num_events)]

    @property
    def y_events(self):
        b = self._b
        if b.num_events == 0:
            return None
        if b.events_packed:
            off = b.event_offsets
            return [np.array(b.event_y[off[self._i * b.num_events + e]:off[self._i * b.num_events + e + 1], :]).T
                    for e in range(b.num_events)]
        return [np.array(b.event_y[self._i, e, :b.event_count[self._i, e], :]).T for e in range(b.num_events)]

    @property
    def event_terminated(self):
        return self._b.num_events > 0 and int(self._b.event_terminate_index[self._i]) >= 0

    @property
    def event_terminate_index(self):
        if self._b.num_events == 0 or int(self._b.event_terminate_index[self._i]) < 0:
            return SIZE_MAX            # the reference's "size_t nan", cysolution.hpp:53
        return int(self._b.event_terminate_index[self._i])

    def __call__(self, t):
        """sol(t): (num_dy, 1) for a scalar, (num_dy, len(t)) for an array (cysolver_api.pyx:96-130)."""
        ts = np.atleast_1d(np.asarray(t, dtype=np.float64))
        out, codes = self._b.call(ts)
        code = int(codes[self._i])
        if code != 0:
            raise CyrkB200Error(_lib.status_message(code))
        return np.array(out[self._i]).T


class BatchResult:
    """One batch solve: per-trajectory arrays + ``result[i]`` views with the WrapCySolverResult attributes.

    The arrays are zero-copy views onto the pinned host mirrors of the native handle (device pointers wrapped in
    :class:`DeviceArray` for device-resident results).  They are created on first access: ``status``, ``steps_taken`` and
    ``y_last`` (and stored rows / events) were copied back by the solve itself, the other per-trajectory arrays
    (``size``, ``n_rejected``, ``nfev``, ``t_last``) are fetched from the device arena when first asked for.
    Lifetime: every view keeps the native handle alive; ``close()`` and passing the result as ``solution_reuse``
    invalidate the result (further attribute access raises) and any view still held by the caller.
    """

    # name -> (getter, dtype, shape function of the result)
    _ARRAYS = {
        "status": ("cyrk_b200_get_status_ptr", np.int32, lambda r: (r.n_traj,)),
        "size": ("cyrk_b200_get_size_ptr", np.uint32, lambda r: (r.n_traj,)),
        "steps_taken": ("cyrk_b200_get_steps_taken_ptr", np.uint32, lambda r: (r.n_traj,)),
        "n_rejected": ("cyrk_b200_get_n_rejected_ptr", np.uint32, lambda r: (r.n_traj,)),
        "nfev": ("cyrk_b200_get_nfev_ptr", np.uint32, lambda r: (r.n_traj,)),
        "t_last": ("cyrk_b200_get_t_last_ptr", np.float64, lambda r: (r.n_traj,)),
        "y_last": ("cyrk_b200_get_y_last_ptr", np.float64, lambda r: (r.n_traj, r.num_dy)),
        "t_rows": ("cyrk_b200_get_t_ptr", np.float64, lambda r: (r.n_traj, r.row_capacity)),
        "y_rows": ("cyrk_b200_get_y_ptr", np.float64, lambda r: (r.n_traj, r.row_capacity, r.num_dy)),
        "event_count": ("cyrk_b200_get_event_count_ptr", np.uint32, lambda r: (r.n_traj, r.num_events)),
        # events_packed: hits of list l = traj * num_events + e are event_t[event_offsets[l]:event_offsets[l + 1]]
        "event_offsets": ("cyrk_b200_get_event_offsets_ptr", np.uint64, lambda r: (r.n_traj * r.num_events + 1,)),
        "event_t": ("cyrk_b200_get_event_t_ptr", np.float64,
                    lambda r: (r.event_total,) if r.events_packed else (r.n_traj, r.num_events, r.event_capacity)),
        "event_y": ("cyrk_b200_get_event_y_ptr", np.float64,
                    lambda r: (r.event_total, r.num_dy) if r.events_packed else (r.n_traj, r.num_events, r.event_capacity, r.num_dy)),
        "event_terminate_index": ("cyrk_b200_get_event_terminate_index_ptr", np.int32, lambda r: (r.n_traj,)),
        "num_interpolates": ("cyrk_b200_get_num_interpolates_ptr", np.uint32, lambda r: (r.n_traj,)),
    }

    def __init__(self, handle, method: int, keep_inputs=(), device_io: bool = False):
        L = _lib.lib()
        self._box = handle if isinstance(handle, _HandleBox) else _HandleBox(handle)
        self._cache = {}
        self._keep = keep_inputs
        self.method = method
        self.device_io = device_io
        h = self._box.h
        self.n_traj = L.cyrk_b200_get_n_traj(h)
        self.num_y = L.cyrk_b200_get_num_y(h)
        self.num_dy = L.cyrk_b200_get_num_dy(h)
        self.row_capacity = L.cyrk_b200_get_row_capacity(h)
        self.num_events = L.cyrk_b200_get_num_events(h)
        self.event_capacity = L.cyrk_b200_get_event_capacity(h) if self.num_events else 0
        self.event_total = L.cyrk_b200_get_event_total(h) if self.num_events else 0
        self.events_packed = bool(self.num_events) and self.event_capacity == 0 and not device_io
        self.kernel_ms = L.cyrk_b200_get_kernel_ms(h)
        self.total_ms = L.cyrk_b200_get_total_ms(h)
        self.launches = L.cyrk_b200_get_launch_count(h)
        self.h2d_bytes = L.cyrk_b200_get_h2d_bytes(h)
        self.d2h_bytes = L.cyrk_b200_get_d2h_bytes(h)

    @property
    def _h(self):
        if self._box is None or not self._box.h:
            raise CyrkB200Error("this BatchResult was closed or handed to solution_reuse: its arrays are gone")
        return self._box.h

    def __getattr__(self, name):
        spec = BatchResult._ARRAYS.get(name)
        if spec is None:
            raise AttributeError(name)
        cache = self.__dict__.get("_cache")
        if cache is None:
            raise AttributeError(name)
        if name in cache:
            return cache[name]
        getter, dtype, shape_of = spec
        h = self._h
        if name in ("t_rows", "y_rows") and not self.row_capacity:
            val = None
        elif name.startswith("event_") and not self.num_events:
            raise AttributeError(f"{name}: the solve had no events")
        else:
            ptr = getattr(_lib.lib(), getter)(h)
            shape = shape_of(self)
            if name == "num_interpolates" and not ptr:
                val = None
            elif not ptr and int(np.prod(shape)) > 0:
                raise CyrkB200Error(_lib.last_error() or f"{getter} returned NULL")
            elif self.device_io:
                val = DeviceArray(ptr, shape, dtype, self._box)
            else:
                val = _np_view(ptr, shape, dtype, self._box)
        cache[name] = val
        return val

    @property
    def lazy_d2h_bytes(self):
        return _lib.lib().cyrk_b200_get_lazy_d2h_bytes(self._h)

    @property
    def success(self):
        return (self.status == 1) | (self.status == 2)

    def __len__(self):
        return self.n_traj

    def __getitem__(self, i) -> TrajectoryView:
        if not -self.n_traj <= i < self.n_traj:
            raise IndexError(i)
        return TrajectoryView(self, i % self.n_traj)

    def call(self, t):
        """Batched sol(t): returns (out[B, len(t), num_dy], codes[B]) -- CySolverResult::call_vectorize
        for every trajectory in one kernel launch."""
        if self.device_io:
            raise CyrkB200Error("call() on a device-resident result: use cyrk_b200_call_vectorize with device pointers")
        ts = np.ascontiguousarray(np.atleast_1d(t), dtype=np.float64)
        out = np.full((self.n_traj, ts.size, self.num_dy), np.nan)
        codes = np.zeros(self.n_traj, dtype=np.int32)
        rc = _lib.lib().cyrk_b200_call_vectorize(self._h, ts.ctypes.data, ts.size, out.ctypes.data, codes.ctypes.data)
        if rc != 0:
            raise CyrkB200Error(_lib.last_error())
        return out, codes

    def ensemble_sums(self):
        """(sum y_last, sum y_last^2, count) over the successful trajectories, reduced on the device in a fixed order
        (cyrk_b200_ensemble_sums).  All-reduce the concatenation across ranks for a multi-GPU ensemble mean/variance
        (cyrk_b200.sharding.ensemble_reduce)."""
        if self.device_io:
            raise CyrkB200Error("device-resident result: call cyrk_b200_ensemble_sums with a device pointer")
        out = np.zeros(2 * self.num_dy + 1)
        if _lib.lib().cyrk_b200_ensemble_sums(self._h, out.ctypes.data) != 0:
            raise CyrkB200Error(_lib.last_error())
        return out[:self.num_dy], out[self.num_dy:2 * self.num_dy], int(out[-1])

    def _release(self):
        """Give up the native handle without freeing it (solution_reuse hands it to the next result)."""
        box, self._box = self._box, None
        self._cache = {}
        return box

    def close(self):
        box = self._release()
        if box is not None:
            box.close()


def cysolve_ivp_batch(
        diffeq,
        t_span,
        y0,
        method: str = "RK45",
        t_eval=None,
        dense_output: bool = False,
        events: Optional[Sequence] = None,
        args=None,
        expected_size: int = 0,
        num_extra: int = 0,
        first_step: float = 0.0,
        max_step: float = INF,
        rtol=1.0e-3,
        atol=1.0e-6,
        max_num_steps: int = 0,
        max_ram_MB: int = 2000,
        store: str = "rows",
        strict_fp: bool = False,
        device: int = 0,
        solution_reuse: Optional[BatchResult] = None,
        devices: Optional[Sequence[int]] = None,
) -> BatchResult:
    """Solve ``B`` independent initial-value problems on one B200.

    Same names/defaults as ``pysolve_ivp`` (pysolver.pyx:348-368) plus:
      y0        (B, n_y); args (n_args,) shared or (B, n_args) per trajectory (float64, the
                reference's ``double[::1] args`` convention) or ``bytes`` for an opaque struct blob;
                t_span (2,) shared or (B, 2); t_eval (n,) shared or (B, n) per trajectory;
                rtol / atol scalar, (n_y,) per component, or (B, n_y) / (B, 1) per trajectory.
      store     'rows'  -> every row the reference would save (``result[i].t/.y``)
                'final' -> only the last row and the counters (``result.y_last``)
      strict_fp run the -fmad=false kernels (the reference's rounding: no FMA contraction)
      devices   several CUDA ordinals: the batch is cut into contiguous shards, one per device, solved concurrently by
                the native library (one host thread + stream per device, no communication); the call then returns
                a list of ``(lo, hi, BatchResult)`` -- shard r holds the trajectories ``[lo, hi)``
    """
    L = _lib.lib()
    method_int = find_ode_method_int(method)
    # pysolver.pyx:255-263
    if first_step < 0.0:
        raise AttributeError("ERROR: `cysolve_ivp_batch` - `first_step` must be a positive number (or 0 for automatic).")
    if max_step <= 0.0:
        raise AttributeError("ERROR: `cysolve_ivp_batch` - `max_step` must be a postive number.")
    rid = _rhs_id(diffeq)

    device_io = _is_cuda_tensor(y0)
    keep = ()
    if device_io:
        # Optional tensor I/O: y0 / args / t_span are contiguous float64 CUDA tensors on `device`;
        # results stay on the device (BatchResult arrays become DeviceArray objects).
        import torch
        y0 = y0.contiguous()
        if y0.dtype != torch.float64 or y0.dim() != 2:
            raise ValueError("device y0 must be a float64 CUDA tensor of shape (B, n_y)")
        B, ny = (int(v) for v in y0.shape)
        device = y0.device.index if y0.device.index is not None else device
        ts_t = t_span if _is_cuda_tensor(t_span) else torch.as_tensor(np.asarray(t_span, dtype=np.float64), device=y0.device)
        ts_t = ts_t.contiguous()
        n_t_span = 1 if ts_t.dim() == 1 else B
        a_t = None
        if args is not None:
            a_t = args if _is_cuda_tensor(args) else torch.as_tensor(np.asarray(args, dtype=np.float64), device=y0.device)
            a_t = a_t.contiguous()
        keep = (y0, ts_t, a_t)
        torch.cuda.current_stream(y0.device).synchronize()   # inputs were produced on torch's stream

        class _P:      # minimal stand-ins exposing .ctypes.data like numpy arrays
            def __init__(self, t): self.ctypes = type("c", (), {"data": t.data_ptr()})
        y0 = _P(y0)
        ts = _P(ts_t)
        if a_t is None:
            a_buf, args_len, n_args = None, 0, 0
        else:
            a_buf = _P(a_t)
            args_len = int(a_t.shape[-1]) * 8
            n_args = 1 if a_t.dim() == 1 else B
    else:
        y0 = np.ascontiguousarray(y0, dtype=np.float64)
        if y0.ndim == 1:
            y0 = y0[None, :]
        if y0.ndim != 2:
            raise ValueError("y0 must have shape (B, n_y)")
        B, ny = y0.shape
        ts = np.ascontiguousarray(t_span, dtype=np.float64)
        if ts.shape not in ((2,), (B, 2)):
            raise ValueError("t_span must have shape (2,) or (B, 2)")
        n_t_span = 1 if ts.ndim == 1 else B

    if device_io:
        pass
    elif args is None:
        a_buf, args_len, n_args = None, 0, 0
    elif isinstance(args, (bytes, bytearray)):
        a_buf = np.frombuffer(bytes(args), dtype=np.uint8).copy()
        args_len, n_args = a_buf.size, 1
    else:
        a = np.ascontiguousarray(args, dtype=np.float64)
        if a.ndim == 1:
            args_len, n_args = a.size * 8, 1
        elif a.ndim == 2 and a.shape[0] == B:
            args_len, n_args = a.shape[1] * 8, B
        else:
            raise ValueError("args must have shape (n_args,) or (B, n_args)")
        a_buf = a
    rt = np.atleast_1d(np.asarray(rtol, dtype=np.float64)).copy()
    at = np.atleast_1d(np.asarray(atol, dtype=np.float64)).copy()
    tol_per_traj = rt.ndim == 2 or at.ndim == 2
    if tol_per_traj:
        # one tolerance row per trajectory: (B, n_y), or (B, 1) broadcast over the components
        try:
            rt = np.ascontiguousarray(np.broadcast_to(rt if rt.ndim == 2 else rt[None, :], (B, ny)))
            at = np.ascontiguousarray(np.broadcast_to(at if at.ndim == 2 else at[None, :], (B, ny)))
        except ValueError:
            raise ValueError("per-trajectory rtol / atol must have shape (B, n_y) or (B, 1)") from None
    te = None if t_eval is None else np.ascontiguousarray(t_eval, dtype=np.float64)
    t_eval_per_traj = te is not None and te.ndim == 2
    if t_eval_per_traj and te.shape[0] != B:
        raise ValueError("per-trajectory t_eval must have shape (B, n_t_eval)")
    n_te = 0 if te is None else (te.shape[1] if t_eval_per_traj else te.size)

    events = list(events) if events else []
    ev_ids = np.zeros(max(len(events), 1), dtype=np.int32)
    ev_dirs = np.zeros(max(len(events), 1), dtype=np.int32)
    ev_max = np.full(max(len(events), 1), SIZE_MAX, dtype=np.uint64)
    for i, ev in enumerate(events):
        if not isinstance(ev, Event):
            ev = Event(*ev) if isinstance(ev, tuple) else Event(ev)
        ev_ids[i] = EVENT_IDS[ev.func] if isinstance(ev.func, str) else int(ev.func)
        ev_dirs[i] = ev.direction
        ev_max[i] = SIZE_MAX if ev.max_allowed is None else int(ev.max_allowed)

    store_mode = {"final": STORE_FINAL, "rows": STORE_ROWS, "steps": STORE_ROWS, "t_eval": STORE_ROWS}[store]
    flags = ((FLAG_STRICT_FP if strict_fp else 0) | (FLAG_DEVICE_IO if device_io else 0) |
             (FLAG_T_EVAL_PER_TRAJ if t_eval_per_traj else 0) | (FLAG_TOL_PER_TRAJ if tol_per_traj else 0))
    call_args = (
        rid, method_int, B,
        ts.ctypes.data, n_t_span,
        y0.ctypes.data, ny,
        int(num_extra),
        a_buf.ctypes.data if a_buf is not None else None, args_len, n_args,
        int(max_num_steps), int(max_ram_MB),
        int(bool(dense_output)),
        te.ctypes.data if te is not None else None, n_te,
        ev_ids.ctypes.data, ev_dirs.ctypes.data, ev_max.ctypes.data, len(events),
        rt.ctypes.data, rt.size,
        at.ctypes.data, at.size,
        float(max_step), float(first_step),
        int(expected_size),
        store_mode, flags, int(device))
    if devices is not None and len(devices) > 0:
        if device_io or solution_reuse is not None:
            raise ValueError("devices=[...] takes host arrays and fresh result handles")
        nd = len(devices)
        dev = (C.c_int * nd)(*[int(d) for d in devices])
        hs = (C.c_void_p * nd)()
        bounds = (C.c_size_t * (nd + 1))()
        if L.cyrk_b200_solve_batch_sharded(dev, nd, hs, bounds, *call_args[:-1]) != 0:
            raise CyrkB200Error(_lib.last_error())
        return [(int(bounds[r]), int(bounds[r + 1]), BatchResult(hs[r], method_int, keep_inputs=keep)) for r in range(nd)]
    if solution_reuse is not None:
        # CySolverResult reuse (Documentation/CySolverResult_Reuses.md): same arena, new problem
        _ = solution_reuse._h                   # raises if that result was already closed / reused
        box = solution_reuse._release()         # the old result (and views taken from it) is invalid from here on
        if L.cyrk_b200_solve_batch_noreturn(box.h, *call_args) != 0:
            box.close()
            raise CyrkB200Error(_lib.last_error())
        return BatchResult(box, method_int, keep_inputs=keep, device_io=device_io)
    h = L.cyrk_b200_solve_batch(*call_args)
    if not h:
        raise CyrkB200Error(_lib.last_error())
    return BatchResult(h, method_int, keep_inputs=keep, device_io=device_io)


def interpolate_from_solution_list(t, solutions: Sequence[BatchResult], breakpoints: Sequence[float]):
    """Batched analogue of ``interpolate_from_solution_list`` (CyRK/cy/helpers.pyx:5-34): several dense batch solutions
    over consecutive spans are evaluated on one time grid.  ``breakpoints[k]`` is the upper time bound of
    ``solutions[k]`` (ascending ``t``); a point is handed to the next solution once it EXCEEDS the current breakpoint,
    exactly like the reference loop.  Returns ``out[B, len(t), num_dy]`` and the per-trajectory call codes."""
    ts = np.ascontiguousarray(np.atleast_1d(t), dtype=np.float64)
    if len(solutions) != len(breakpoints) or not solutions:
        raise ValueError("one breakpoint per solution is required")
    B, ndy = solutions[0].n_traj, solutions[0].num_dy
    out = np.full((B, ts.size, ndy), np.nan)
    codes = np.zeros(B, dtype=np.int32)
    owner = np.zeros(ts.size, dtype=np.int64)
    k = 0
    for i, x in enumerate(ts):                       # same sequential hand-over rule as the reference
        while x > breakpoints[k] and k + 1 < len(solutions):
            k += 1
        owner[i] = k
    for k, sol in enumerate(solutions):
        sel = np.nonzero(owner == k)[0]
        if sel.size == 0:
            continue
        o, c = sol.call(ts[sel])
        out[:, sel, :] = o
        codes = np.where(codes != 0, codes, c)
    return out, codes
