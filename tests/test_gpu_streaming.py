"""GPU: the streamed (chunked, double-buffered) solve and the compact wire formats give the SAME results as the
single-arena call -- bit for bit, because every chunk runs the unmodified single-batch code on its slice.

Covers VERDICT r1 item 3 (rows / t_eval samples / event lists streamed chunk by chunk, dense_output kept on the device
as step records) -- reference storage being replaced: CySolverResult::save_data / record_event_data
(CyRK/cy/cysolution.cpp:396-448), reuse contract CyRK/cy/cysolve.cpp:6-59.
"""
import numpy as np
import pytest

import cyrk_b200 as c

pytestmark = pytest.mark.gpu


def _opt(name, value):
    assert c.lib().cyrk_b200_set_option(name.encode(), int(value)) == 0


@pytest.fixture
def streaming():
    """stream_chunk forces chunking of small test batches; always restored."""
    yield _opt
    _opt("stream", 1); _opt("stream_chunk", 0); _opt("dense_store", 0)


def _both(streaming, chunk, **kw):
    streaming("stream", 0)
    a = c.cysolve_ivp_batch(**kw)
    streaming("stream", 1); streaming("stream_chunk", chunk)
    b = c.cysolve_ivp_batch(**kw)
    streaming("stream_chunk", 0)
    return a, b


def _same_scalars(a, b, names=("status", "size", "steps_taken", "n_rejected", "nfev", "t_last", "y_last")):
    for name in names:
        assert np.array_equal(np.asarray(getattr(a, name)), np.asarray(getattr(b, name)), equal_nan=True), name


def _same_rows(a, b):
    for i in range(a.n_traj):
        n = int(a.size[i])
        assert np.array_equal(a.t_rows[i, :n], b.t_rows[i, :n]), i
        assert np.array_equal(a.y_rows[i, :n], b.y_rows[i, :n]), i


@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fma"])
def test_streamed_teval_rows_and_dense_records_equal_the_single_arena_call(streaming, strict):
    """BASELINE config 2's feature mix (Lorenz DOP853, t_eval + dense_output) in 5 chunks of 1024 (last one short)."""
    rng = np.random.default_rng(5)
    B = 4500
    args = np.stack([np.full(B, 10.0), rng.uniform(20, 35, B), np.full(B, 8 / 3)], 1)
    y0 = np.tile([1.0, 0.0, 0.0], (B, 1))
    te = np.linspace(0.0, 4.0, 97)
    kw = dict(diffeq="lorenz", t_span=(0.0, 4.0), y0=y0, method="DOP853", args=args, rtol=1e-9, atol=1e-11,
              t_eval=te, dense_output=True, strict_fp=strict, expected_size=160)
    a, b = _both(streaming, 1024, **kw)
    assert (a.status == 1).all() and b.launches > a.launches
    _same_scalars(a, b)
    _same_rows(a, b)
    assert np.array_equal(a.num_interpolates, b.num_interpolates)
    tq = np.sort(rng.uniform(0.0, 4.0, 33))
    (qa, ca), (qb, cb) = a.call(tq), b.call(tq)
    assert np.array_equal(ca, cb) and np.array_equal(qa, qb)
    # sol(t) at the t_eval times reproduces the emitted samples (same interpolant, same evaluation)
    qs, _ = b.call(te[1:-1])
    assert np.allclose(qs, np.asarray(b.y_rows)[:, 1:-1, :], rtol=1e-13, atol=1e-13)


def test_streamed_step_rows_replan_when_a_later_chunk_needs_more_capacity(streaming):
    """store='rows' without t_eval: the probe chunk fixes the row capacity; trajectories of later chunks integrate 3x
    longer and overflow it, so the whole batch is re-planned -- the result must still equal the single-arena call."""
    rng = np.random.default_rng(6)
    B = 3000
    y0 = np.stack([rng.uniform(5, 15, B), rng.uniform(2.5, 7.5, B)], 1)
    ts = np.stack([np.zeros(B), np.where(np.arange(B) < 1000, 3.0, 10.0)], 1)
    kw = dict(diffeq="lotkavolterra", t_span=ts, y0=y0, method="RK45", args=[1.5, 1, 3, 1], rtol=1e-7, atol=1e-9,
              store="rows", expected_size=40)
    a, b = _both(streaming, 1000, **kw)
    assert (a.status == 1).all()
    _same_scalars(a, b)
    _same_rows(a, b)
    # and with the interpolants kept (step-parallel path: the rows ARE the step records)
    kw["dense_output"] = True
    a, b = _both(streaming, 1000, **kw)
    _same_scalars(a, b)
    _same_rows(a, b)
    tq = np.linspace(0.0, 3.0, 17)
    (qa, ca), (qb, cb) = a.call(tq), b.call(tq)
    assert np.array_equal(ca, cb) and np.array_equal(qa, qb)


@pytest.mark.parametrize("store", ["final", "rows"])
def test_streamed_events_with_extras_equal_the_single_arena_call(streaming, store):
    """BASELINE config 3's feature mix: damped pendulum RK23, 1 extra output, three events (one terminal at a
    per-trajectory time).  Packed event lists are stitched across chunks with global offsets."""
    rng = np.random.default_rng(7)
    B = 2500
    y0 = np.stack([rng.uniform(-1.5, 1.5, B), rng.uniform(-1, 1, B)], 1)
    args = np.stack([np.ones(B), np.ones(B), np.full(B, 9.81), np.full(B, 1.5), rng.uniform(2.0, 12.0, B)], 1)
    ev = [c.Event("y0"), c.Event("y1", direction=-1), c.Event("t_minus_arg4", terminal=True)]
    kw = dict(diffeq="pendulum_drag", t_span=(0.0, 10.0), y0=y0, method="RK23", args=args, rtol=1e-6, atol=1e-8,
              events=ev, num_extra=1, store=store, expected_size=2048 if store == "rows" else 0)
    a, b = _both(streaming, 512, **kw)
    assert {1, 2} <= set(np.unique(a.status).tolist())
    _same_scalars(a, b)
    assert a.events_packed and b.events_packed and a.event_total == b.event_total
    assert np.array_equal(a.event_count, b.event_count) and np.array_equal(a.event_offsets, b.event_offsets)
    assert np.array_equal(a.event_t, b.event_t) and np.array_equal(a.event_y, b.event_y)
    assert np.array_equal(a.event_terminate_index, b.event_terminate_index)
    # offsets are the running sum of the counts, list by list
    cnt = np.asarray(a.event_count, dtype=np.uint64).ravel()
    assert np.array_equal(np.asarray(a.event_offsets), np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint64))
    for i in (0, 511, 512, 1700, B - 1):
        va, vb = a[i], b[i]
        assert [len(t) for t in va.t_events] == [int(n) for n in a.event_count[i]]
        for ta, tb in zip(va.t_events, vb.t_events):
            assert np.array_equal(ta, tb)
        for ya, yb in zip(va.y_events, vb.y_events):
            assert ya.shape[0] == 3 and np.array_equal(ya, yb)
    if store == "rows":
        _same_rows(a, b)


def test_dense_with_events_is_not_streamed_and_still_correct(streaming):
    """A dense store that cannot be kept as step records (events / extras) must outlive its chunk: such a call stays on
    the single-arena path even when chunking is forced."""
    y0 = np.tile([0.0, 1.0], (600, 1))
    kw = dict(diffeq="accuracy", t_span=(0.0, 10.0), y0=y0, method="RK45", rtol=1e-8, atol=1e-9, dense_output=True,
              events=[c.Event("y0")])
    a, b = _both(streaming, 128, **kw)
    assert a.launches == b.launches
    _same_scalars(a, b)
    t = np.linspace(0.1, 9.9, 11)
    assert np.array_equal(a.call(t)[0], b.call(t)[0])


@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fma"])
@pytest.mark.parametrize("method", ["RK23", "RK45", "DOP853"])
def test_step_records_as_dense_store_equal_materialised_interpolants(streaming, method, strict):
    """dense_store = 0 (default): sol(t) recomputes the queried step from its 40-byte record; dense_store = 1: every
    [y_old | Q] record is written (cysolution.cpp:450-482) and read back.  Same instructions on the same inputs."""
    rng = np.random.default_rng(8)
    B = 400
    y0 = np.stack([rng.uniform(5, 15, B), rng.uniform(2.5, 7.5, B)], 1)
    tq = np.sort(rng.uniform(0.0, 10.0, 50))
    outs = []
    for store in (0, 1):
        streaming("dense_store", store)
        for span in ((0.0, 10.0), (10.0, 0.0)):
            r = c.cysolve_ivp_batch("lotkavolterra", span, y0, method=method, args=[1.5, 1, 3, 1], rtol=1e-7, atol=1e-9,
                                    dense_output=True, strict_fp=strict)
            assert (r.status == 1).all()
            outs.append((np.array(r.num_interpolates), np.array(r.nfev)) + r.call(tq))
    for (na, fa, qa, ca), (nb, fb, qb, cb) in zip(outs[:2], outs[2:]):
        assert np.array_equal(na, nb) and np.array_equal(fa, fb) and np.array_equal(ca, cb)
        assert np.array_equal(qa, qb, equal_nan=True)
