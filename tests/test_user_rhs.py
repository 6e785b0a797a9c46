"""User right-hand sides compiled with NVRTC behind the same C ABI (cyrk_b200_rhs_compile)."""
import numpy as np
import pytest

import cyrk_b200 as c

LV_SRC = r'''
// same signature family as the reference's DiffeqFuncType (c_common.hpp:218-220)
__device__ void my_lv(double* dy, double t, const double* y, const char* args) {
    const double* p = (const double*)args;
    const double a = p[0], b = p[1], cc = p[2], d = p[3];
    const double y0 = y[0], y1 = y[1];
    dy[0] = a * y0 - b * y0 * y1;
    dy[1] = -cc * y1 + d * y0 * y1;
}
'''

DRAG_SRC = r'''
struct PendArgs { double l, m, g, c, stop; };          // opaque struct blob, Advanced_CySolver.md:7-13
__device__ double torque_of(double t) { return 0.1 * sin(t); }     // a "pre-eval" helper is just another function
__device__ void drag(double* dy, double t, const double* y, const char* args) {
    const PendArgs* p = (const PendArgs*)args;
    const double coeff_1 = (-3. * p->g / (2. * p->l));
    const double coeff_2 = (3. / (p->m * (p->l * p->l)));
    const double torque = torque_of(t);
    dy[0] = y[1];
    double d1 = coeff_1 * sin(y[0]) + coeff_2 * torque;
    d1 -= p->c * y[1];
    dy[1] = d1;
    dy[2] = torque;
}
__device__ double drag_event(int id, double t, const double* y, const char* args) {
    const PendArgs* p = (const PendArgs*)args;
    if (id == 0) return y[0];
    if (id == 1) return y[1];
    return t - p->stop;
}
'''


def test_compile_cache_and_error_log():
    u = c.UserRHS(LV_SRC, "my_lv", n_y=2, args_bytes=32)
    assert u.rhs_id >= 1000
    assert c.UserRHS(LV_SRC, "my_lv", n_y=2, args_bytes=32).rhs_id == u.rhs_id      # cached by source hash
    import ctypes as C
    a, b, d = C.c_size_t(), C.c_size_t(), C.c_size_t()
    assert c.lib().cyrk_b200_rhs_info(u.rhs_id, C.byref(a), C.byref(b), C.byref(d)) == 0
    assert (a.value, b.value, d.value) == (2, 0, 32)
    with pytest.raises(c.CyrkB200Error, match="undefined"):
        c.UserRHS("__device__ void bad(double* dy, double t, const double* y, const char* a) { dy[0] = nope; }", "bad", n_y=1)
    assert c.UserRHS(LV_SRC, "my_lv", n_y=64, args_bytes=32).rhs_id >= 1000   # whole-vector form above 32: CTA path, one thread evaluates it
    with pytest.raises(c.CyrkB200Error, match="component-wise"):
        c.UserRHS(LV_SRC, "my_lv", n_y=64, n_extra=1, args_bytes=32)


@pytest.mark.gpu
def test_user_lv_equals_builtin():
    u = c.UserRHS(LV_SRC, "my_lv", n_y=2, args_bytes=32)
    rng = np.random.default_rng(0)
    B = 4096
    y0 = np.stack([rng.uniform(5, 15, B), rng.uniform(2.5, 7.5, B)], 1)
    for method in ("RK23", "RK45", "DOP853"):
        kw = dict(method=method, args=[1.5, 1, 3, 1], rtol=1e-6, atol=1e-8)
        a = c.cysolve_ivp_batch(u, (0, 10), y0, strict_fp=True, **kw)
        b = c.cysolve_ivp_batch("lotkavolterra", (0, 10), y0, strict_fp=True, **kw)
        assert np.array_equal(a.status, b.status) and np.array_equal(a.steps_taken, b.steps_taken)
        assert np.array_equal(a.y_last, b.y_last)                  # -fmad=false: NVRTC and nvcc agree bit for bit
        a = c.cysolve_ivp_batch(u, (0, 10), y0, **kw)
        b = c.cysolve_ivp_batch("lotkavolterra", (0, 10), y0, **kw)
        assert np.array_equal(a.status, b.status) and (a.steps_taken == b.steps_taken).mean() >= 0.999
        assert np.all(np.abs(a.y_last - b.y_last) <= 10 * 1e-6 * np.abs(b.y_last) + 10 * 1e-8)


@pytest.mark.gpu
def test_user_rhs_dense_and_t_eval_take_the_step_parallel_path():
    """NVRTC right-hand side + dense_output + t_eval (no events / extras): VAR_STEPS solve + step_dense_body compiled at run
    time.  Strict build: bitwise equal to the precompiled built-in; both: bitwise equal to the per-trajectory kernel."""
    u = c.UserRHS(LV_SRC, "my_lv", n_y=2, args_bytes=32)
    rng = np.random.default_rng(5)
    B = 300
    y0 = np.stack([rng.uniform(5, 15, B), rng.uniform(2.5, 7.5, B)], 1)
    te = np.linspace(0, 10, 57)
    tq = np.linspace(0.5, 9.5, 9)
    for method in ("RK45", "DOP853"):
        kw = dict(method=method, args=[1.5, 1, 3, 1], rtol=1e-6, atol=1e-8, t_eval=te, dense_output=True)
        a = c.cysolve_ivp_batch(u, (0, 10), y0, strict_fp=True, **kw)
        b = c.cysolve_ivp_batch("lotkavolterra", (0, 10), y0, strict_fp=True, **kw)
        assert a.launches == b.launches >= 3                      # recording solve + step_dense + last row
        assert np.array_equal(a.size, b.size) and np.array_equal(a.nfev, b.nfev)
        assert np.array_equal(a.y_rows, b.y_rows) and np.array_equal(a.call(tq)[0], b.call(tq)[0])
        try:
            c.lib().cyrk_b200_set_option(b"two_pass", 0)
            d = c.cysolve_ivp_batch(u, (0, 10), y0, **kw)
        finally:
            c.lib().cyrk_b200_set_option(b"two_pass", 1)
        e = c.cysolve_ivp_batch(u, (0, 10), y0, **kw)
        assert d.launches < e.launches
        assert np.array_equal(d.y_rows, e.y_rows) and np.array_equal(d.nfev, e.nfev) and np.array_equal(d.call(tq)[0], e.call(tq)[0])


@pytest.mark.gpu
def test_user_rhs_with_struct_args_events_extras_dense(oracle):
    u = c.UserRHS(DRAG_SRC, "drag", n_y=2, n_extra=1, args_bytes=40, n_events=3)
    rng = np.random.default_rng(1)
    B = 64
    y0 = np.stack([rng.uniform(-1.5, 1.5, B), rng.uniform(-1, 1, B)], 1)
    args = np.stack([np.ones(B), np.ones(B), np.full(B, 9.81), np.full(B, 1.5), rng.uniform(2.0, 12.0, B)], 1)
    ev = [c.Event(0), c.Event(1, direction=-1), c.Event(2, terminal=True)]
    r = c.cysolve_ivp_batch(u, (0.0, 10.0), y0, method="RK45", args=args, rtol=1e-6, atol=1e-8, events=ev, num_extra=1,
                            dense_output=True)
    for i in range(0, B, 7):
        ref = oracle.solve("pendulum_drag", (0.0, 10.0), y0[i], "RK45", 1e-6, 1e-8, args=args[i], dense_output=True,
                           events=[("y0", 0, None), ("y1", -1, None), ("t_minus_arg4", 0, 1)], fix_quirk11=True,
                           call_t=[0.1, 0.7, 1.5])
        v = r[i]
        assert (v.status, v.size, v.steps_taken) == (ref.status, ref.size, ref.steps_taken)
        assert [len(t) for t in v.t_events] == [len(t) for t in ref.t_events]
        assert np.all(np.abs(v.y - ref.y) <= 10 * 1e-6 * np.abs(ref.y) + 10 * 1e-8)
        assert np.allclose(v(np.array([0.1, 0.7, 1.5])), ref.dense, rtol=1e-5, atol=1e-7)


CHAIN_SRC = r'''
// component-wise form for LARGE systems (CTA-per-trajectory path): returns dy_i
__device__ double chain(unsigned i, double t, const double* y, unsigned n, const char* args) {
    const double r = ((const double*)args)[0];
    const double left = i > 0 ? y[i - 1] : 0.0, right = i + 1 < n ? y[i + 1] : 0.0;
    return r * (left - 2.0 * y[i] + right) + 0.1 * sin(t);
}
'''


CHAIN_X_SRC = r'''
// the same chain with two extra outputs (CyRK's dy[n_y:], cysolver_test.pyx:42-57): the functor answers i = n and n + 1
__device__ double chainx(unsigned i, double t, const double* y, unsigned n, const char* args) {
    const double r = ((const double*)args)[0];
    if (i == n) return y[0] * y[n - 1] + t;               // extra 0
    if (i == n + 1) return r * (y[1] - y[0]);             // extra 1
    const double left = i > 0 ? y[i - 1] : 0.0, right = i + 1 < n ? y[i + 1] : 0.0;
    return r * (left - 2.0 * y[i] + right) + 0.1 * sin(t);
}
'''


def test_large_user_rhs_compiles_on_cpu():
    u = c.UserRHS(CHAIN_SRC, "chain", n_y=0, args_bytes=8)
    assert u.rhs_id >= 1000
    ux = c.UserRHS(CHAIN_X_SRC, "chainx", n_y=0, n_extra=2, args_bytes=8)      # components n, n+1 are extra outputs
    assert ux.rhs_id >= 1000 and ux.rhs_id != u.rhs_id


@pytest.mark.gpu
@pytest.mark.parametrize("method,n", [("RK23", 96), ("RK45", 96), ("DOP853", 96), ("RK45", 700), ("DOP853", 1100)])
def test_large_user_rhs_matches_oracle(oracle, method, n):
    """A user-supplied large system (discrete diffusion chain; n_y = 96 / 700 / 1100 -> 1 / 2 / 4 components per
    thread of the register-resident CTA kernel) through NVRTC vs the CPU oracle driven with the same right-hand side
    as a ctypes callback."""
    import ctypes as C
    B = 3
    u = c.UserRHS(CHAIN_SRC, "chain", n_y=0, args_bytes=8)
    rng = np.random.default_rng(9)
    y0 = rng.uniform(0.0, 1.0, (B, n))
    rates = np.array([[0.5], [0.8], [1.1]])

    @oracle.DIFFEQ_FN
    def chain_cb(dy, t, y, args, pre):
        yy = np.ctypeslib.as_array(y, shape=(n,))
        r = C.cast(args, C.POINTER(C.c_double))[0]
        left = np.concatenate([[0.0], yy[:-1]])
        right = np.concatenate([yy[1:], [0.0]])
        out = r * (left - 2.0 * yy + right) + 0.1 * np.sin(t)
        C.memmove(dy, out.ctypes.data, 8 * n)

    te = np.linspace(0.0, 2.0, 7)
    for t_eval in (None, te):
        r = c.cysolve_ivp_batch(u, (0.0, 2.0), y0, method=method, args=rates, rtol=1e-6, atol=1e-9, t_eval=t_eval)
        for i in range(B):
            ref = oracle.solve(chain_cb, (0.0, 2.0), y0[i], method, 1e-6, 1e-9, args=rates[i], t_eval=t_eval, num_extra=0)
            v = r[i]
            assert (v.status, v.size, v.steps_taken) == (ref.status, ref.size, ref.steps_taken)
            assert np.allclose(v.t, ref.t, rtol=1e-6, atol=1e-12)
            assert np.all(np.abs(v.y - ref.y) <= 10 * 1e-6 * np.abs(ref.y) + 10 * 1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("method,n", [("RK23", 96), ("RK45", 700), ("DOP853", 1100), ("RK45", 5000)])
def test_large_user_rhs_with_extra_outputs(oracle, method, n):
    """Extra outputs on the CTA-per-trajectory path (register-resident kernel for n_y <= 4096, slab kernel above): every
    saved row carries dy[n_y:] of the right-hand-side call at that row (cysolver.cpp:968-972 save_data), rows are
    n_y + num_extra wide like CySolverResult::solution."""
    import ctypes as C
    B = 3
    u = c.UserRHS(CHAIN_X_SRC, "chainx", n_y=0, n_extra=2, args_bytes=8)
    rng = np.random.default_rng(29)
    y0 = rng.uniform(0.0, 1.0, (B, n))
    rates = np.array([[0.5], [0.8], [1.1]])

    @oracle.DIFFEQ_FN
    def cb(dy, t, y, args, pre):
        yy = np.ctypeslib.as_array(y, shape=(n,))
        r = C.cast(args, C.POINTER(C.c_double))[0]
        left = np.concatenate([[0.0], yy[:-1]])
        right = np.concatenate([yy[1:], [0.0]])
        out = np.concatenate([r * (left - 2.0 * yy + right) + 0.1 * np.sin(t), [yy[0] * yy[n - 1] + t, r * (yy[1] - yy[0])]])
        C.memmove(dy, out.ctypes.data, 8 * (n + 2))

    for store in ("rows", "final"):
        r = c.cysolve_ivp_batch(u, (0.0, 2.0), y0, method=method, args=rates, rtol=1e-6, atol=1e-9, num_extra=2, store=store)
        assert r.num_y == n and r.num_dy == n + 2
        for i in range(B):
            ref = oracle.solve(cb, (0.0, 2.0), y0[i], method, 1e-6, 1e-9, args=rates[i], num_extra=2)
            assert (int(r.status[i]), int(r.size[i]), int(r.steps_taken[i]), int(r.nfev[i])) == (ref.status, ref.size, ref.steps_taken, ref.nfev)
            tol = 10 * 1e-6 * np.abs(ref.y) + 10 * 1e-9
            assert np.all(np.abs(r.y_last[i] - ref.y[:, -1]) <= tol[:, -1] + 1e-12)
            if store == "rows":
                v = r[i]
                assert v.y.shape == (n + 2, ref.size) and np.allclose(v.t, ref.t, rtol=1e-6, atol=1e-12)
                assert np.all(np.abs(v.y - ref.y) <= tol + 1e-9)
    with pytest.raises(c.CyrkB200Error, match="raw step rows / final state only"):
        c.cysolve_ivp_batch(u, (0.0, 2.0), y0, method=method, args=rates, num_extra=2, t_eval=np.linspace(0, 2, 5))


CHAIN_EV_SRC = CHAIN_SRC + r'''
// events of a large user system: g(id, t, y[n], n, args)
__device__ double chain_event(int id, double t, const double* y, unsigned n, const char* args) {
    if (id == 0) return y[0];
    if (id == 1) return y[1];
    return (t > 5.0) ? 0.0 : 1.0;
}
'''


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["RK45", "DOP853"])
def test_large_user_rhs_with_events(oracle, method):
    """Large user system (NVRTC, CTA path) with user event functions: zero crossings of y[0] (any direction) and y[1]
    (falling), and a terminal step-function event at t = 5; vs the oracle with the same right-hand side as a callback and
    the built-in events of the same meaning."""
    import ctypes as C
    n, B = 200, 3
    u = c.UserRHS(CHAIN_EV_SRC, "chain", n_y=0, args_bytes=8, n_events=3)
    rng = np.random.default_rng(19)
    y0 = rng.uniform(-0.5, 0.5, (B, n))
    rates = np.array([[0.5], [0.8], [1.1]])

    @oracle.DIFFEQ_FN
    def chain_cb(dy, t, y, args, pre):
        yy = np.ctypeslib.as_array(y, shape=(n,))
        r = C.cast(args, C.POINTER(C.c_double))[0]
        left = np.concatenate([[0.0], yy[:-1]])
        right = np.concatenate([yy[1:], [0.0]])
        out = r * (left - 2.0 * yy + right) + 0.1 * np.sin(t)
        C.memmove(dy, out.ctypes.data, 8 * n)

    ev = [c.Event(0), c.Event(1, direction=-1), c.Event(2, terminal=True)]
    r = c.cysolve_ivp_batch(u, (0.0, 8.0), y0, method=method, args=rates, rtol=1e-6, atol=1e-9, events=ev)
    for i in range(B):
        ref = oracle.solve(chain_cb, (0.0, 8.0), y0[i], method, 1e-6, 1e-9, args=rates[i], num_extra=0,
                           events=[("y0", 0, None), ("y1", -1, None), ("t_gt_5", 0, 1)])
        v = r[i]
        assert (v.status, v.size, v.steps_taken) == (ref.status, ref.size, ref.steps_taken) and v.status == 2
        assert [len(t) for t in v.t_events] == [len(t) for t in ref.t_events]
        for a, b in zip(v.t_events, ref.t_events):
            assert np.allclose(a, b, rtol=1e-6, atol=1e-9)
        for a, b in zip(v.y_events, ref.y_events):
            assert np.all(np.abs(a - b) <= 10 * 1e-6 * np.abs(b) + 10 * 1e-9)
        assert v.event_terminate_index == ref.event_terminate_index == 2
        assert np.all(np.abs(v.y - ref.y) <= 10 * 1e-6 * np.abs(ref.y) + 10 * 1e-9)
