"""GPU: events with store='final' run as a detection solve plus one thread per triggered step (VAR_EVDETECT +
event_roots_kernel, rk_core.cuh).  Redoing a triggered step from (t_old, t_now, step, y_old) executes the same
instructions on the same inputs as the one-pass kernel, so EVERYTHING must agree bit for bit: counters, nfev, terminal
state, hit lists, termination index.  Reference: p_check_events (CyRK/cy/cysolver.cpp:563-750), c_brentq
(optimize/c_brentq.cpp:41-204), termination (cysolver.cpp:814-834)."""
import numpy as np
import pytest

import cyrk_b200 as c

pytestmark = pytest.mark.gpu


def _opt(name, value):
    assert c.lib().cyrk_b200_set_option(name.encode(), int(value)) == 0


def _both(**kw):
    try:
        _opt("events_two_pass", 0)
        a = c.cysolve_ivp_batch(**kw)
        _opt("events_two_pass", 1)
        b = c.cysolve_ivp_batch(**kw)
    finally:
        _opt("events_two_pass", 1)
    return a, b


def _same(a, b):
    for name in ("status", "size", "steps_taken", "n_rejected", "nfev", "t_last", "y_last", "event_count", "event_offsets",
                 "event_t", "event_y", "event_terminate_index"):
        assert np.array_equal(np.asarray(getattr(a, name)), np.asarray(getattr(b, name)), equal_nan=True), name


@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fma"])
@pytest.mark.parametrize("method", ["RK23", "RK45", "DOP853"])
def test_c3_mix_two_pass_equals_one_pass(method, strict):
    rng = np.random.default_rng(50)
    B = 3000
    y0 = np.stack([rng.uniform(-1.5, 1.5, B), rng.uniform(-1, 1, B)], 1)
    args = np.stack([np.ones(B), np.ones(B), np.full(B, 9.81), np.full(B, 1.5), rng.uniform(2.0, 12.0, B)], 1)
    ev = [c.Event("y0"), c.Event("y1", direction=-1), c.Event("t_minus_arg4", terminal=True)]
    a, b = _both(diffeq="pendulum_drag", t_span=(0.0, 10.0), y0=y0, method=method, args=args, rtol=1e-6, atol=1e-8,
                 events=ev, num_extra=1, store="final", strict_fp=strict)
    assert {1, 2} <= set(np.unique(a.status).tolist())
    assert b.launches == a.launches + 1              # detection solve + event_roots_kernel
    _same(a, b)


@pytest.mark.parametrize("backward", [False, True], ids=["fwd", "bwd"])
def test_max_allowed_direction_filters_and_step_function_events(backward):
    """max_allowed > 1 (termination on the n-th hit, later hits of the same step dropped), rising / falling filters,
    the reference's step-function events (0 -> 0 triggers on every step once active, quirk 4), no extras."""
    rng = np.random.default_rng(51)
    B = 1500
    y0 = np.stack([rng.uniform(-1.5, 1.5, B), rng.uniform(-1, 1, B)], 1)
    span = (8.0, 0.0) if backward else (0.0, 8.0)
    for ev in ([c.Event("y0", max_allowed=3), c.Event("y1", direction=1), c.Event("y1", direction=-1, max_allowed=5)],
               [c.Event("t_gt_5"), c.Event("y0", direction=-1, max_allowed=2)],
               [c.Event("y0_gt_50"), c.Event("y1_lt_50", max_allowed=40)]):
        a, b = _both(diffeq="pendulum", t_span=span, y0=y0, method="RK45", args=[1.0, 1.0, 9.81], rtol=1e-7, atol=1e-9,
                     events=ev, store="final", expected_size=1200)
        _same(a, b)
    assert (np.asarray(a.status) == 2).any()


def test_two_pass_events_match_the_oracle(oracle):
    rng = np.random.default_rng(52)
    B = 200
    y0 = np.stack([rng.uniform(-1.5, 1.5, B), rng.uniform(-1, 1, B)], 1)
    args = np.stack([np.ones(B), np.ones(B), np.full(B, 9.81), np.full(B, 1.5), rng.uniform(2.0, 12.0, B)], 1)
    ev = [c.Event("y0"), c.Event("y1", direction=-1), c.Event("t_minus_arg4", terminal=True)]
    r = c.cysolve_ivp_batch("pendulum_drag", (0.0, 10.0), y0, method="RK23", args=args, rtol=1e-6, atol=1e-8,
                            events=ev, num_extra=1, store="final", strict_fp=True)
    assert r.launches >= 2
    for i in range(0, B, 3):
        ref = oracle.solve("pendulum_drag", (0.0, 10.0), y0[i], "RK23", 1e-6, 1e-8, args=args[i],
                           events=[("y0", 0, None), ("y1", -1, None), ("t_minus_arg4", 0, 1)], fix_quirk11=True)
        v = r[i]
        assert (v.status, v.size, v.steps_taken) == (ref.status, ref.size, ref.steps_taken)
        assert [len(t) for t in v.t_events] == [len(t) for t in ref.t_events]
        for ta, tb in zip(v.t_events, ref.t_events):
            assert np.allclose(ta, tb, rtol=1e-5, atol=1e-9)
        for ya, yb in zip(v.y_events, ref.y_events):
            assert np.all(np.abs(ya - yb) <= 10 * 1e-6 * np.abs(yb) + 10 * 1e-8 + 1e-7)
        assert v.event_terminated == ref.event_terminated
        if ref.event_terminated:
            assert v.event_terminate_index == ref.event_terminate_index
        assert np.all(np.abs(r.y_last[i] - ref.y[:, -1]) <= 10 * 1e-6 * np.abs(ref.y[:, -1]) + 10 * 1e-8)
        assert abs(r.t_last[i] - ref.t[-1]) <= 1e-5 * abs(ref.t[-1])


def test_event_queue_overflow_reruns_with_the_exact_capacity():
    """More triggered steps than the queue's first guess (16 per trajectory): a step-function event triggers on EVERY
    step once it is active (quirk 4), the pendulum takes ~400 steps."""
    y0 = np.tile([0.3, 0.0], (64, 1))
    a, b = _both(diffeq="pendulum", t_span=(0.0, 10.0), y0=y0, method="RK45", args=[1.0, 1.0, 9.81], rtol=1e-8, atol=1e-10,
                 events=[c.Event("t_gt_5"), c.Event("y0")], store="final", expected_size=2000)
    assert int(np.asarray(a.event_count)[:, 0].min()) > 16
    _same(a, b)


def test_user_rhs_events_two_pass(oracle):
    from test_user_rhs import DRAG_SRC
    u = c.UserRHS(DRAG_SRC, "drag", n_y=2, n_extra=1, args_bytes=40, n_events=3)
    rng = np.random.default_rng(53)
    B = 500
    y0 = np.stack([rng.uniform(-1.5, 1.5, B), rng.uniform(-1, 1, B)], 1)
    args = np.stack([np.ones(B), np.ones(B), np.full(B, 9.81), np.full(B, 1.5), rng.uniform(2.0, 12.0, B)], 1)
    ev = [c.Event(0), c.Event(1, direction=-1), c.Event(2, terminal=True)]
    a, b = _both(diffeq=u, t_span=(0.0, 10.0), y0=y0, method="RK45", args=args, rtol=1e-6, atol=1e-8, events=ev,
                 num_extra=1, store="final")
    _same(a, b)
