"""Randomised configuration sweeps (fixed seeds), modelled on the reference's big parameter sweep
(Tests/E_CySolver_Tests/test_a_cysolve_ivp.py:21-126):

* CPU: C restatement == compiled reference bit for bit on random feature mixes (when oracle/_ref is present);
* GPU: CUDA path vs the oracle on random feature mixes -- methods x direction x t_eval x dense x events x extras x
  scalar/array tolerances x first_step/max_step -- to the parity bar of BASELINE.json.
"""
import numpy as np
import pytest

import cyrk_b200 as c

PROBLEMS = {
    "lotkavolterra": dict(ny=2, nx=0, args=lambda r: [r.uniform(1.2, 1.8), 1.0, 3.0, 1.0], y0=lambda r: [r.uniform(5, 15), r.uniform(2.5, 7.5)]),
    "extraoutput": dict(ny=2, nx=2, args=lambda r: None, y0=lambda r: [r.uniform(15, 25), r.uniform(15, 25)]),
    "pendulum_drag": dict(ny=2, nx=1, args=lambda r: [1.0, 1.0, 9.81, r.uniform(0.5, 2.0), r.uniform(1.0, 9.0)],
                          y0=lambda r: [r.uniform(-1.5, 1.5), r.uniform(-1, 1)]),
    "accuracy": dict(ny=2, nx=0, args=lambda r: None, y0=lambda r: [r.uniform(-1, 1), r.uniform(-1, 1)]),
    "lorenz_extraoutput": dict(ny=3, nx=3, args=lambda r: [10.0, r.uniform(20, 30), 8 / 3], y0=lambda r: [1.0, 0.0, 0.0]),
}
EVENT_POOL = [("y0", 0, None), ("y1", -1, None), ("y1", 1, 3), ("t_gt_5", 0, None), ("y0", 1, 2), ("t_minus_arg4", 0, 1)]


def random_config(seed):
    r = np.random.default_rng(seed)
    name = list(PROBLEMS)[r.integers(len(PROBLEMS))]
    P = PROBLEMS[name]
    method = ["rk23", "rk45", "dop853"][r.integers(3)]
    t1 = float(r.uniform(2.0, 8.0))
    backward = bool(r.integers(4) == 0) and name in ("lotkavolterra", "extraoutput", "accuracy")
    span = (t1, 0.0) if backward else (0.0, t1)
    kw = dict(rhs=name, t_span=span, y0=P["y0"](r), args=P["args"](r), method=method)
    if r.integers(2):
        kw["rtol"], kw["atol"] = 10.0 ** r.uniform(-8, -4), 10.0 ** r.uniform(-10, -6)
    else:
        kw["rtol"] = (10.0 ** r.uniform(-8, -4, P["ny"])).tolist()
        kw["atol"] = (10.0 ** r.uniform(-10, -6, P["ny"])).tolist()
    if method == "rk23":                              # keep RK23 step counts small
        kw["rtol"] = np.maximum(kw["rtol"], 1e-6).tolist() if isinstance(kw["rtol"], list) else max(kw["rtol"], 1e-6)
    mode = r.integers(4)
    if mode in (1, 3):
        npts = int(r.integers(2, 30))
        hi = t1 * (1.0 if r.integers(2) else r.uniform(0.3, 0.9))
        kw["t_eval"] = (np.linspace(t1, t1 - hi, npts) if backward else np.linspace(0.0, hi, npts))
    if mode in (2, 3):
        kw["dense_output"] = True
        kw["call_t"] = np.sort(r.uniform(0.0, t1, 6))
    if r.integers(3) == 0 and not backward:
        k = int(r.integers(1, 4))
        pool = [e for e in EVENT_POOL if e[0] != "t_minus_arg4" or name == "pendulum_drag"]
        idx = r.choice(len(pool), size=min(k, len(pool)), replace=False)
        kw["events"] = [pool[i] for i in idx]
    knob = r.integers(5)
    if knob == 0:
        kw["first_step"] = float(10.0 ** r.uniform(-4, -2))
    elif knob == 1:
        kw["max_step"] = float(r.uniform(0.05, 0.5))
    elif knob == 2:
        kw["max_num_steps"] = int(r.integers(3, 40))
    return kw


@pytest.mark.parametrize("seed", range(120))
def test_port_equals_reference_on_random_configs(oracle, seed):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built on this box")
    kw = random_config(seed)
    a = oracle.solve(impl="ref", **kw)
    b = oracle.solve(impl="port", **kw)
    assert (a.status, a.size, a.steps_taken, a.nfev, a.num_interpolates) == (b.status, b.size, b.steps_taken, b.nfev, b.num_interpolates)
    assert np.array_equal(a.t, b.t) and np.array_equal(a.y, b.y, equal_nan=True)
    assert a.event_terminated == b.event_terminated and a.event_terminate_index == b.event_terminate_index
    for x, y in zip(a.t_events, b.t_events):
        assert np.array_equal(x, y)
    for x, y in zip(a.y_events, b.y_events):
        assert np.array_equal(x, y, equal_nan=True)
    if a.dense is not None:
        assert a.dense_code == b.dense_code and np.array_equal(a.dense, b.dense, equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fma"])
@pytest.mark.parametrize("seed", range(150))
def test_gpu_equals_oracle_on_random_configs(oracle, seed, strict):
    kw = random_config(seed)
    ref = oracle.solve(impl="port", fix_quirk11=True, **kw)
    g = dict(kw)
    call_t = g.pop("call_t", None)
    rhs = g.pop("rhs"); y0 = np.asarray(g.pop("y0"), dtype=np.float64)[None, :]; span = g.pop("t_span")
    evs = [c.Event(n, direction=d, max_allowed=m) for (n, d, m) in g.pop("events", [])]
    r = c.cysolve_ivp_batch(rhs, span, y0, events=evs, num_extra=PROBLEMS[rhs]["nx"], strict_fp=strict, store="rows", **g)
    v = r[0]
    rt, at = float(np.max(kw["rtol"])), float(np.max(kw["atol"]))
    assert v.status == ref.status, (kw, v.status, ref.status)
    assert (v.size, v.steps_taken) == (ref.size, ref.steps_taken), kw
    if evs and PROBLEMS[rhs]["nx"] > 0:
        # with extra outputs every Brent iteration costs one RHS evaluation, and the iteration count of a root search
        # may legitimately differ by one between two FP64 implementations
        hits = sum(len(t) for t in ref.t_events)
        assert abs(int(r.nfev[0]) - ref.nfev) <= 2 * max(hits, 1), kw
    else:
        assert int(r.nfev[0]) == ref.nfev, kw
    # accepted-t sequence within 1e-6 relative (to the time scale of the problem)
    assert np.allclose(v.t, ref.t, rtol=1e-6, atol=1e-6 * abs(span[1] - span[0]))
    slack = 0.0
    if kw.get("t_eval") is None and ref.t.size > 2:
        with np.errstate(all="ignore"):
            dydt = np.nan_to_num(np.abs(np.gradient(ref.y, ref.t, axis=1)), nan=0.0, posinf=0.0, neginf=0.0)
        slack = 2.0 * dydt * np.abs(v.t - ref.t)[None, :]
    assert np.all(np.abs(v.y - ref.y) <= 10 * rt * np.abs(ref.y) + 10 * at + slack + 1e-9 * ("lorenz" in rhs)), kw
    if evs:
        assert v.event_terminated == ref.event_terminated and v.event_terminate_index == ref.event_terminate_index
        for a, b in zip(v.t_events, ref.t_events):
            assert len(a) == len(b) and np.allclose(a, b, rtol=10 * rt, atol=1e-9)
        for a, b in zip(v.y_events, ref.y_events):
            assert np.all(np.abs(a - b) <= 10 * rt * np.abs(b) + 10 * at + 1e-7)
    if call_t is not None:
        out, codes = r.call(call_t)
        assert int(codes[0]) == ref.dense_code
        if ref.dense_code == 0:
            ok = ~np.isnan(ref.dense)
            assert np.all(np.abs(out[0].T[ok] - ref.dense[ok]) <= 10 * rt * np.abs(ref.dense[ok]) + 10 * at + 1e-7), kw
