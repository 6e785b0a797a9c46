"""CPU: the C restatement (oracle port) against (i) the committed golden vectors generated from the compiled
reference and (ii), where oracle/_ref is present, the compiled reference itself -- bit for bit."""
import numpy as np
import pytest

from cases import cases

CASES = cases()


def _check(r, g, name, exact=True):
    sc = g[f"{name}/scalars"]
    assert r.status == sc[0], name
    assert r.size == sc[1] and r.steps_taken == sc[2] and r.num_interpolates == sc[3], name
    assert r.nfev == sc[4], name
    assert int(r.event_terminated) == sc[5], name
    ti = -1 if r.event_terminate_index == (1 << 64) - 1 else r.event_terminate_index
    assert ti == sc[6], name
    assert np.array_equal(r.t, g[f"{name}/t"]), name
    assert np.array_equal(r.y, g[f"{name}/y"], equal_nan=True), name
    for e in range(len(r.t_events)):
        assert np.array_equal(r.t_events[e], g[f"{name}/t_events{e}"]), name
        assert np.array_equal(r.y_events[e], g[f"{name}/y_events{e}"], equal_nan=True), name
    if r.dense is not None:
        assert r.dense_code == g[f"{name}/dense_code"][0], name
        assert np.array_equal(r.dense, g[f"{name}/dense"], equal_nan=True), name


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_matches_golden_bitwise(oracle, golden, name):
    r = oracle.solve(impl="port", **CASES[name])
    _check(r, golden, name)


def test_golden_covers_status_codes(golden):
    st = {int(golden[k][0]) for k in golden.files if k.endswith("/scalars")}
    assert {1, 2, -52, -53, -54} <= st


@pytest.mark.parametrize("name", sorted(CASES)[::5])
def test_port_matches_compiled_reference(oracle, name):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built on this box")
    a = oracle.solve(impl="ref", **CASES[name])
    b = oracle.solve(impl="port", **CASES[name])
    assert (a.status, a.size, a.steps_taken, a.nfev) == (b.status, b.size, b.steps_taken, b.nfev)
    assert np.array_equal(a.t, b.t) and np.array_equal(a.y, b.y, equal_nan=True)
    for x, y in zip(a.t_events, b.t_events):
        assert np.array_equal(x, y)


def test_port_vs_reference_random_batch(oracle):
    """prange-style batch: random ICs, port == compiled reference bit for bit, any thread count."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built on this box")
    rng = np.random.default_rng(0)
    y0 = np.stack([rng.uniform(5, 15, 300), rng.uniform(2.5, 7.5, 300)], 1)
    kw = dict(rhs="lotkavolterra", t_span=(0, 10), y0=y0, method="RK45", rtol=1e-7, atol=1e-9, args=[1.5, 1, 3, 1])
    a = oracle.solve_batch(impl="ref", n_threads=1, **kw)
    b = oracle.solve_batch(impl="port", n_threads=3, **kw)
    c = oracle.solve_batch(impl="ref", n_threads=4, **kw)
    for k in ("status", "size", "steps", "nfev", "t_last", "y_last"):
        assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], c[k])


@pytest.mark.parametrize("n", [96, 700])
def test_port_matches_reference_largey_tolerance_arrays_and_t_eval(oracle, n):
    """What tests/test_gpu_largey.py::test_cta_per_trajectory_tolerances_and_t_eval leans on: the port equals the compiled
    reference bit for bit for a large system with per-component rtol / atol arrays, with and without t_eval, forward and
    backward (the reversed-t_eval copy of cysolver.cpp:520-527)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built on this box")
    rng = np.random.default_rng(7 + n)
    y0 = 100.0 * (1 + 0.01 * rng.uniform(-1, 1, n))
    rt = 10.0 ** rng.uniform(-8, -5) * np.where(np.arange(n) % 3 == 0, 0.5, 1.0)
    at = rt * 1e-2
    args = [-0.47, float(n)]
    for span in ((0.0, 10.0), (10.0, 0.0)):
        te = np.sort(rng.uniform(-0.5, 10.0, 7))
        te[-1] = 10.0
        if span[0] > span[1]:
            te = te[::-1].copy()
        for t_eval in (None, te):
            a = oracle.solve("largey", span, y0, "RK45", rt, at, args=args, t_eval=t_eval, impl="ref")
            b = oracle.solve("largey", span, y0, "RK45", rt, at, args=args, t_eval=t_eval, impl="port")
            assert (a.status, a.size, a.steps_taken, a.nfev) == (b.status, b.size, b.steps_taken, b.nfev), (n, span)
            assert np.array_equal(a.t, b.t) and np.array_equal(a.y, b.y, equal_nan=True)
