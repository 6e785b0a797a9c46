"""GPU: the CTA-per-trajectory path (large n_y) against the oracle and the golden vectors."""
import numpy as np
import pytest

import cyrk_b200 as c
from cases import cases

pytestmark = pytest.mark.gpu
CASES = cases()


def _tol(ref, rtol, atol):
    return 10 * rtol * np.abs(ref) + 10 * atol


@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fma"])
@pytest.mark.parametrize("name", ["largey_rk45_n96", "largey_simple_rk45_n64"])
def test_golden_large(golden, name, strict):
    kw = dict(CASES[name])
    rhs = kw.pop("rhs"); y0 = np.asarray(kw.pop("y0"))[None, :]; ts = kw.pop("t_span")
    r = c.cysolve_ivp_batch(rhs, ts, y0, strict_fp=strict, **kw)
    sc = golden[f"{name}/scalars"]
    v = r[0]
    assert (v.status, v.size, v.steps_taken) == (sc[0], sc[1], sc[2])
    assert int(r.nfev[0]) == sc[4]
    assert np.allclose(v.t, golden[f"{name}/t"], rtol=1e-6, atol=1e-12)
    yr = golden[f"{name}/y"]
    assert np.all(np.abs(v.y - yr) <= _tol(yr, kw["rtol"], kw["atol"]))


@pytest.mark.parametrize("method", ["RK23", "RK45", "DOP853"])
@pytest.mark.parametrize("n", [37, 300])
def test_large_methods_and_t_eval(oracle, method, n):
    rng = np.random.default_rng(5)
    B = 6
    y0 = 100.0 * (1 + 0.01 * rng.uniform(-1, 1, (B, n)))
    args = np.stack([rng.uniform(-0.6, -0.4, B), np.full(B, float(n))], 1)       # per-trajectory decay rate
    te = np.linspace(0, 10, 9)
    for t_eval in (None, te):
        r = c.cysolve_ivp_batch("largey", (0.0, 10.0), y0, method=method, args=args, rtol=1e-6, atol=1e-8, t_eval=t_eval)
        for i in range(B):
            ref = oracle.solve("largey", (0.0, 10.0), y0[i], method, 1e-6, 1e-8, args=args[i], t_eval=t_eval)
            v = r[i]
            assert (v.status, v.size, v.steps_taken) == (ref.status, ref.size, ref.steps_taken), (method, n, i)
            assert int(r.nfev[i]) == ref.nfev
            assert np.allclose(v.t, ref.t, rtol=1e-6, atol=1e-12)      # parity metric: accepted-t sequence within 1e-6 relative
            assert np.all(np.abs(v.y - ref.y) <= _tol(ref.y, 1e-6, 1e-8))
            assert np.all(np.abs(r.y_last[i] - ref.y[:, -1]) <= _tol(ref.y[:, -1], 1e-6, 1e-8))


def test_c4_largey_4096(oracle):
    """BASELINE config 4 shape (n_y = 4096, RK45, rtol 1e-6 / atol 1e-8) on a few trajectories vs the oracle."""
    rng = np.random.default_rng(3)
    B, n = 6, 4096
    y0 = 100.0 * (1 + 0.01 * rng.uniform(-1, 1, (B, n)))
    r = c.cysolve_ivp_batch("largey", (0.0, 10.0), y0, method="RK45", args=[-0.5, float(n)], rtol=1e-6, atol=1e-8)
    ob = oracle.solve_batch("largey", (0.0, 10.0), y0, "RK45", 1e-6, 1e-8, args=[-0.5, float(n)], impl="port")
    assert np.array_equal(r.status, ob["status"]) and (r.status == 1).all()
    assert np.array_equal(r.steps_taken, ob["steps"]) and np.array_equal(r.size, ob["size"])
    assert np.all(np.abs(r.y_last - ob["y_last"]) <= _tol(ob["y_last"], 1e-6, 1e-8))
    # rows are complete and time-ordered
    for i in range(B):
        assert np.all(np.diff(r[i].t) > 0) and r[i].t[-1] == 10.0 and r[i].y.shape == (n, r[i].size)


def test_largey_simple_and_unsupported_features():
    n = 2048
    y0 = 100.0 * np.ones((3, n))
    r = c.cysolve_ivp_batch("largey_simple", (0.0, 10.0), y0, method="RK45", args=[-0.5, float(n)], rtol=1e-6, atol=1e-8, store="final")
    assert (r.status == 1).all()
    # dy_0 = sin(2 pi t / 10), the rest is constant: y_0(10) = 100 + (10/2pi)(1 - cos(2 pi)) = 100
    assert np.allclose(r.y_last[:, 0], 100.0, atol=5e-3) and np.all(r.y_last[:, 1:] == 100.0)
    with pytest.raises(c.CyrkB200Error):        # no extra outputs on the CTA path
        c.cysolve_ivp_batch("largey", (0.0, 10.0), y0, args=[-0.5, float(n)], num_extra=1)
    # stored dense output / events need the register-resident kernel: n_y <= 4096 for DOP853, n_y <= 12 288 for RK23 / RK45
    big = 100.0 * np.ones((2, 5000))
    with pytest.raises(c.CyrkB200Error):
        c.cysolve_ivp_batch("largey", (0.0, 10.0), big, method="DOP853", args=[-0.5, 5000.0], dense_output=True)
    with pytest.raises(c.CyrkB200Error):
        c.cysolve_ivp_batch("largey", (0.0, 10.0), big, method="DOP853", args=[-0.5, 5000.0], events=[c.Event("y0")])
    huge = 100.0 * np.ones((2, 12500))
    with pytest.raises(c.CyrkB200Error):
        c.cysolve_ivp_batch("largey", (0.0, 10.0), huge, args=[-0.5, 12500.0], events=[c.Event("y0")])


@pytest.mark.parametrize("n", [700, 1500, 3000, 5000, 8192, 10000, 12000, 12500])
def test_register_resident_kernel_sizes(oracle, n):
    """n_y = 700 / 1500 / 3000 exercise C = 2 / 4 / 8 components per thread with a ragged last block of components
    (cta2_core.cuh, every vector on chip); 5000 / 8192 / 10 000 / 12 000 take C = 12 / 16 / 20 / 24 with the stage
    derivatives (and from C = 20 the aux table) in the L2-resident slab -- RK23 / RK45; DOP853 at those sizes, and
    n_y = 12 500, run on the shared-memory/slab kernel (cta_core.cuh)."""
    rng = np.random.default_rng(11)
    B = 3
    y0 = 100.0 * (1 + 0.01 * rng.uniform(-1, 1, (B, n)))
    args = [-0.5, float(n)]
    te = np.linspace(0, 10, 5)
    cases = (("RK45", None), ("RK23", te), ("DOP853", te)) if n <= 5000 else (("RK45", None), ("RK45", te), ("RK23", te))
    for method, t_eval in cases:
        ref = oracle.solve("largey", (0.0, 10.0), y0[1], method, 1e-6, 1e-8, args=args, t_eval=t_eval)
        for strict in (False, True):
            r = c.cysolve_ivp_batch("largey", (0.0, 10.0), y0, method=method, args=args, rtol=1e-6, atol=1e-8, t_eval=t_eval, strict_fp=strict)
            v = r[1]
            assert (v.status, v.size, v.steps_taken) == (ref.status, ref.size, ref.steps_taken), (method, n, strict)
            assert int(r.nfev[1]) == ref.nfev
            assert np.allclose(v.t, ref.t, rtol=1e-6, atol=1e-12)
            assert np.all(np.abs(v.y - ref.y) <= _tol(ref.y, 1e-6, 1e-8))


def test_large_slab_range_agrees_with_the_slab_kernel():
    """4096 < n_y <= 12 288 on the register-resident kernel (option cta2_large, default) vs the shared-memory/slab kernel."""
    rng = np.random.default_rng(13)
    n, B = 10000, 5
    y0 = 100.0 * (1 + 0.01 * rng.uniform(-1, 1, (B, n)))
    out = []
    for large in (1, 0):
        c.lib().cyrk_b200_set_option(b"cta2_large", large)
        try:
            out.append(c.cysolve_ivp_batch("largey", (0.0, 10.0), y0, method="RK45", args=[-0.5, float(n)], rtol=1e-6, atol=1e-8))
        finally:
            c.lib().cyrk_b200_set_option(b"cta2_large", 1)
    a, b = out
    assert (a.status == 1).all() and np.array_equal(a.steps_taken, b.steps_taken) and np.array_equal(a.n_rejected, b.n_rejected)
    assert np.array_equal(a.size, b.size) and np.allclose(a.y_last, b.y_last, rtol=1e-9, atol=1e-12)
    for i in range(B):
        assert np.allclose(a[i].y, b[i].y, rtol=1e-9, atol=1e-12) and np.allclose(a[i].t, b[i].t, rtol=1e-10, atol=0)


def test_both_cta_kernels_agree():
    """The register-resident kernel and the shared-memory/slab kernel are two schedules of the same arithmetic:
    same step counts, states equal to rounding (the error-norm partial sums are grouped differently)."""
    rng = np.random.default_rng(12)
    n, B = 1024, 4
    y0 = 100.0 * (1 + 0.01 * rng.uniform(-1, 1, (B, n)))
    out = []
    for kernel in (0, 1):
        c.lib().cyrk_b200_set_option(b"cta_kernel", kernel)
        try:
            out.append(c.cysolve_ivp_batch("largey", (0.0, 10.0), y0, method="RK45", args=[-0.5, float(n)], rtol=1e-6, atol=1e-8))
        finally:
            c.lib().cyrk_b200_set_option(b"cta_kernel", 0)
    a, b = out
    assert np.array_equal(a.steps_taken, b.steps_taken) and np.array_equal(a.n_rejected, b.n_rejected)
    assert np.allclose(a.y_last, b.y_last, rtol=1e-9, atol=1e-12)


def test_cta_per_component_tolerances():
    """Non-uniform rtol/atol arrays take the array path of the register-resident kernel."""
    rng = np.random.default_rng(13)
    n, B = 600, 2
    y0 = 100.0 * (1 + 0.01 * rng.uniform(-1, 1, (B, n)))
    rt = np.full(n, 1e-6); rt[::2] = 1e-7
    at = np.full(n, 1e-8)
    outs = []
    for kernel in (0, 1):
        c.lib().cyrk_b200_set_option(b"cta_kernel", kernel)
        try:
            outs.append(c.cysolve_ivp_batch("largey", (0.0, 10.0), y0, method="RK45", args=[-0.5, float(n)], rtol=rt, atol=at))
        finally:
            c.lib().cyrk_b200_set_option(b"cta_kernel", 0)
    assert np.array_equal(outs[0].steps_taken, outs[1].steps_taken)
    assert np.allclose(outs[0].y_last, outs[1].y_last, rtol=1e-9, atol=1e-12)


def test_cta_results_are_bitwise_reproducible():
    """The register-resident kernel synchronises with one barrier per evaluation and double-buffered shared memory;
    a missing barrier would show up as run-to-run differences.  Same batch three times (more trajectories than CTAs,
    so trajectories land on different CTAs in different orders): bitwise identical rows."""
    rng = np.random.default_rng(21)
    n, B = 4096, 400
    y0 = 100.0 * (1 + 0.01 * rng.uniform(-1, 1, (B, n)))
    ref = None
    for rep in range(3):
        r = c.cysolve_ivp_batch("largey", (0.0, 3.0), y0, method="RK45", args=[-0.5, float(n)], rtol=1e-6, atol=1e-8,
                                t_eval=np.linspace(0.0, 3.0, 4))
        cur = (np.array(r.steps_taken), np.array(r.n_rejected), np.array(r.y_last), np.array(r[B - 1].y), np.array(r[0].y))
        if ref is None:
            ref = cur
        else:
            assert all(np.array_equal(a, b) for a, b in zip(ref, cur))
        r.close()


@pytest.mark.parametrize("method,n", [(m, n) for m in ("RK23", "RK45", "DOP853") for n in (130, 1500)] + [("RK45", 5000), ("RK23", 10000)])
def test_cta_dense_output_matches_oracle(oracle, method, n):
    """(n_y = 5000 / 10 000: the slab-extended range of the register-resident kernel, C = 12 / 20.)  dense_output=True on the CTA path: sol(t) from the stored interpolants vs the oracle's CySolverResult::call, forward
    and backward, with and without t_eval; nfev includes DOP853's three extra evaluations per interpolant."""
    rng = np.random.default_rng(31)
    B = 3
    y0 = 100.0 * (1 + 0.01 * rng.uniform(-1, 1, (B, n)))
    args = [-0.5, float(n)]
    for span, t_eval in (((0.0, 6.0), None), ((6.0, 0.0), None), ((0.0, 6.0), np.linspace(0.0, 6.0, 5))):
        tq = np.array([0.0, 0.37, 2.5, 5.99, 6.0])
        r = c.cysolve_ivp_batch("largey", span, y0, method=method, args=args, rtol=1e-6, atol=1e-8, dense_output=True,
                                t_eval=t_eval, expected_size=4)          # tiny capacity: exercises the exact-capacity re-run
        out, codes = r.call(tq)
        assert (codes == 0).all() and out.shape == (B, tq.size, n)
        for i in (0, B - 1):
            ref = oracle.solve("largey", span, y0[i], method, 1e-6, 1e-8, args=args, dense_output=True, t_eval=t_eval, call_t=tq)
            v = r[i]
            assert (v.status, v.size, v.steps_taken) == (ref.status, ref.size, ref.steps_taken)
            assert int(r.nfev[i]) == ref.nfev
            got = out[i].T                                           # (n, len(tq)) like sol(t)
            # (backward runs: the reference's lookup lands on the degenerate interpolant #0 at t = t_start -> NaN; kept)
            nan = np.isnan(ref.dense)
            assert np.array_equal(np.isnan(got), nan)
            assert np.all(np.abs(got - ref.dense)[~nan] <= _tol(ref.dense, 1e-6, 1e-8)[~nan]), (method, n, span)
            assert np.allclose(v(tq), got, equal_nan=True)
        # outside the span -> ARGUMENT_ERROR (-21... the reference's code), per trajectory
        _, bad = r.call(np.array([7.0]))
        assert (bad != 0).all()
        r.close()


def test_cta_failure_statuses_match_oracle(oracle):
    """Failure paths of the CTA kernels end with the reference's status codes (and never hang): max_num_steps (-52),
    a NaN initial state (step size collapses: -54), a bad first_step (-80, clean instead of the reference's crash)."""
    rng = np.random.default_rng(41)
    n = 700
    y0 = 100.0 * (1 + 0.01 * rng.uniform(-1, 1, (3, n)))
    args = [-0.5, float(n)]
    for kernel in (0, 1):
        c.lib().cyrk_b200_set_option(b"cta_kernel", kernel)
        try:
            # user cap on the number of steps
            r = c.cysolve_ivp_batch("largey", (0.0, 10.0), y0, method="RK45", args=args, rtol=1e-6, atol=1e-8, max_num_steps=5)
            ref = oracle.solve("largey", (0.0, 10.0), y0[0], "RK45", 1e-6, 1e-8, args=args, max_num_steps=5)
            assert ref.status == -52 and (r.status == ref.status).all() and (r.steps_taken == ref.steps_taken).all()
            # NaN in the initial state of ONE trajectory: that one fails, its neighbours are untouched
            bad = y0.copy(); bad[1, 17] = np.nan
            r = c.cysolve_ivp_batch("largey", (0.0, 10.0), bad, method="RK45", args=args, rtol=1e-6, atol=1e-8)
            ref = oracle.solve("largey", (0.0, 10.0), bad[1], "RK45", 1e-6, 1e-8, args=args)
            good = oracle.solve("largey", (0.0, 10.0), bad[0], "RK45", 1e-6, 1e-8, args=args)
            assert r.status[1] == ref.status and ref.status < 0
            assert r.status[0] == 1 and r.status[2] == 1 and r.steps_taken[0] == good.steps_taken
            # first_step larger than half the span
            r = c.cysolve_ivp_batch("largey", (0.0, 10.0), y0, method="RK45", args=args, rtol=1e-6, atol=1e-8, first_step=6.0)
            assert (r.status == -80).all() and (r.steps_taken == 0).all()
        finally:
            c.lib().cyrk_b200_set_option(b"cta_kernel", 0)


@pytest.mark.parametrize("method,n", [(m, n) for m in ("RK23", "RK45", "DOP853") for n in (130, 1500)] + [("RK23", 5000), ("RK45", 10000)])
def test_cta_events_match_oracle(oracle, method, n):
    """(n_y = 5000 / 10 000: the slab-extended range of the register-resident kernel, C = 12 / 20.)  Events on the CTA path (register-resident kernel): the reference's built-in event functions on largey --
    step functions, so every search runs the reference's special cases (f(xa) == 0, f(xb) == 0, bisection on a
    discontinuity) --, non-terminal and terminal, with t_eval and stored dense output, vs the oracle."""
    rng = np.random.default_rng(51)
    B = 3
    y0 = 100.0 * (1 + 0.01 * rng.uniform(-1, 1, (B, n)))
    args = [-0.5, float(n)]
    cases = [
        ([("y0_gt_50", 0, None), ("y1_lt_50", 0, None), ("t_gt_5", 0, None)], None, False),
        ([("t_gt_5", 0, 1)], None, False),                                   # terminal: ends inside a step
        ([("y1_lt_50", 0, None), ("t_gt_5", 0, 3)], np.linspace(0.0, 10.0, 7), True),
        ([("t_gt_5", -1, None), ("t_gt_5", 1, None)], None, False),          # direction filters
    ]
    for evs, t_eval, dense in cases:
        events = [c.Event(name, direction=d, max_allowed=m) for (name, d, m) in evs]
        r = c.cysolve_ivp_batch("largey", (0.0, 10.0), y0, method=method, args=args, rtol=1e-6, atol=1e-8, events=events,
                                t_eval=t_eval, dense_output=dense, expected_size=2)     # tiny capacities: exact re-run
        for i in (0, B - 1):
            ref = oracle.solve("largey", (0.0, 10.0), y0[i], method, 1e-6, 1e-8, args=args, events=evs, t_eval=t_eval,
                               dense_output=dense)
            v = r[i]
            assert (v.status, v.size, v.steps_taken) == (ref.status, ref.size, ref.steps_taken), (method, n, evs)
            assert int(r.nfev[i]) == ref.nfev
            assert [len(t) for t in v.t_events] == [len(t) for t in ref.t_events]
            for a, b in zip(v.t_events, ref.t_events):
                assert np.allclose(a, b, rtol=1e-6, atol=1e-9)
            for a, b in zip(v.y_events, ref.y_events):
                assert a.shape == b.shape and np.all(np.abs(a - b) <= _tol(b, 1e-6, 1e-8))
            assert v.event_terminated == ref.event_terminated
            if ref.event_terminated:
                assert v.event_terminate_index == ref.event_terminate_index
            assert np.allclose(v.t, ref.t, rtol=1e-6, atol=1e-12)
            assert np.all(np.abs(v.y - ref.y) <= _tol(ref.y, 1e-6, 1e-8))
        r.close()


@pytest.mark.parametrize("n,kernel", [(96, 0), (700, 0), (5000, 0), (300, 1)], ids=["n96", "n700", "n5000", "n300-kernel-forced"])
def test_cta_per_trajectory_tolerances_and_t_eval(oracle, n, kernel):
    """SURVEY 8b row 1 on the CTA-per-trajectory path: one (rtol, atol) row and one t_eval row per trajectory, each trajectory
    against ITS OWN oracle solve; both CTA kernels; forward and backward runs."""
    rng = np.random.default_rng(41 + n)
    B, m = 5, 7
    y0 = 100.0 * (1 + 0.01 * rng.uniform(-1, 1, (B, n)))
    args = np.stack([rng.uniform(-0.6, -0.4, B), np.full(B, float(n))], 1)
    rt = 10.0 ** rng.uniform(-8, -5, (B, 1)) * np.where(np.arange(n) % 3 == 0, 0.5, 1.0)[None, :]     # (B, n): rows differ, components too
    at = rt * 1e-2
    c.lib().cyrk_b200_set_option(b"cta_kernel", kernel)
    try:
        # per-trajectory tolerances, raw step rows
        r = c.cysolve_ivp_batch("largey", (0.0, 10.0), y0, method="RK45", args=args, rtol=rt, atol=at)
        for i in range(B):
            ref = oracle.solve("largey", (0.0, 10.0), y0[i], "RK45", rt[i], at[i], args=args[i])
            v = r[i]
            assert (v.status, v.size, v.steps_taken) == (ref.status, ref.size, ref.steps_taken), (n, i)
            assert int(r.nfev[i]) == ref.nfev
            assert np.allclose(v.t, ref.t, rtol=1e-6, atol=1e-12)
            assert np.all(np.abs(v.y - ref.y) <= 10 * rt[i][:, None] * np.abs(ref.y) + 10 * at[i][:, None])
        assert len(set(int(s) for s in r.steps_taken)) > 1          # the rows really were used per trajectory
        # (B, 1) tolerances broadcast over the components == the same scalar pair per trajectory
        r1 = c.cysolve_ivp_batch("largey", (0.0, 10.0), y0, method="RK23", args=args, rtol=rt[:, :1] * 100, atol=at[:, :1] * 100, store="final")
        for i in range(B):
            ref = oracle.solve("largey", (0.0, 10.0), y0[i], "RK23", float(rt[i, 0] * 100), float(at[i, 0] * 100), args=args[i])
            assert (int(r1.status[i]), int(r1.steps_taken[i]), int(r1.nfev[i])) == (ref.status, ref.steps_taken, ref.nfev)
        # per-trajectory t_eval (with per-trajectory tolerances on top), forward and backward
        for span in ((0.0, 10.0), (10.0, 0.0)):
            lo = rng.uniform(-0.5, 3.0, B); hi = rng.uniform(6.0, 10.0, B)
            te = lo[:, None] + (hi - lo)[:, None] * np.sort(rng.uniform(0, 1, (B, m)), axis=1)
            te[0, -1] = 10.0
            if span[0] > span[1]:
                te = te[:, ::-1].copy()
            r2 = c.cysolve_ivp_batch("largey", span, y0, method="RK45", args=args, rtol=rt, atol=at, t_eval=te)
            for i in range(B):
                ref = oracle.solve("largey", span, y0[i], "RK45", rt[i], at[i], args=args[i], t_eval=te[i])
                v = r2[i]
                assert (v.status, v.size, v.steps_taken) == (ref.status, ref.size, ref.steps_taken), (n, span, i)
                assert np.array_equal(v.t, ref.t)
                assert np.all(np.abs(v.y - ref.y) <= 10 * rt[i][:, None] * np.abs(ref.y) + 10 * at[i][:, None])
    finally:
        c.lib().cyrk_b200_set_option(b"cta_kernel", 0)
    # a shared row passed per trajectory is the shared solve of the same kernel, bit for bit (per-trajectory rows run on the
    # shared-memory/slab kernel: option cta_kernel = 1 selects it for the shared solve too)
    c.lib().cyrk_b200_set_option(b"cta_kernel", 1)
    try:
        a = c.cysolve_ivp_batch("largey", (0.0, 10.0), y0, method="RK45", args=args, rtol=rt[0], atol=at[0], store="final")
    finally:
        c.lib().cyrk_b200_set_option(b"cta_kernel", 0)
    b = c.cysolve_ivp_batch("largey", (0.0, 10.0), y0, method="RK45", args=args, rtol=np.tile(rt[0], (B, 1)), atol=np.tile(at[0], (B, 1)), store="final")
    for name in ("status", "steps_taken", "n_rejected", "nfev", "y_last"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    with pytest.raises(c.CyrkB200Error, match="cannot be combined"):
        c.cysolve_ivp_batch("largey", (0.0, 10.0), y0, method="RK45", args=args, rtol=rt, atol=at, events=[c.Event("y0")])
