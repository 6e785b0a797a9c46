"""CPU: the oracle against the reference's OWN known-answer tests and the survey-time known answers."""
import math

import numpy as np
import pytest


def test_dense_extra_kat(oracle):
    """Tests/D_PySolver_Tests/test_b_pysolve_dense.py:41-50 (8-digit constants, np.allclose)."""
    r = oracle.solve("dense_kat", (0.0, 10.0), [5.0, 2.0], "RK45", 1e-3, 1e-6, max_step=100_000, dense_output=True,
                     call_t=[0.3, 4.0, 8.0, 9.10, 9.9])
    assert r.success
    expected = np.asarray([[4.52627329, 1.87160845, 0.13355624, 1.57871591, -1.01590877],
                           [2.13093483, 21.05384256, 10.31201731, 48.55169176, 57.06848508],
                           [0.25, 0.25, 0.25, 0.25, 0.25],
                           [0.15, 2., 4., 4.55, 4.95]])
    assert r.dense.shape == (4, 5)
    assert np.allclose(r.dense, expected)


@pytest.mark.parametrize("i", range(10))
def test_brentq_kat(oracle, i):
    """Tests/F_Optimize_Tests/test_brentq.py: roots of x^3 - a, xtol = rtol = 1e-3, 10 iterations."""
    a0 = -2.0 - i / 10.0
    root, err, _, _ = oracle.brentq_poly3([a0, 0.0, 0.0, 1.0], 0.0, 2.0, 0.001, 0.001, 10)
    assert math.isclose(root, (-a0) ** (1.0 / 3.0), rel_tol=0.001, abs_tol=0.001)
    if oracle.have_ref():
        r2 = oracle.brentq_poly3([a0, 0.0, 0.0, 1.0], 0.0, 2.0, 0.001, 0.001, 10, impl="ref")
        assert r2[0] == root and r2[1] == err


@pytest.mark.parametrize("method,rt,at", [("rk23", 1e-4, 1e-7), ("rk45", 1e-4, 1e-7), ("dop853", 1e-5, 1e-8)])
@pytest.mark.parametrize("t_eval_end", [None, 0.5, 1.0])
def test_closed_form_accuracy(oracle, method, rt, at, t_eval_end):
    """Tests/E_CySolver_Tests/test_a_cysolve_ivp.py:129-195."""
    def exact(t, c1=1.0, c2=0.5):
        return np.stack([-c1 * np.sin(t) + c2 * np.cos(t) - np.cos(t) / 2, c2 * np.sin(t) + c1 * np.cos(t) + np.sin(t) / 2])
    te = None if t_eval_end is None else np.linspace(0.0, t_eval_end, 50)
    tq = np.linspace(0.1, 0.4, 10)
    r = oracle.solve("accuracy", (0.0, 10.0), [0.0, 1.0], method, 1e-8, 1e-9, t_eval=te, dense_output=True, call_t=tq)
    assert r.success
    assert np.allclose(r.y, exact(r.t), rtol=rt, atol=at)
    assert r.dense.shape == (2, 10)
    assert np.allclose(r.dense, exact(tq), rtol=rt, atol=at)


def test_status_code_pins(oracle):
    """Tests/D_PySolver_Tests/test_a_pysolve_ivp.py:349-380."""
    kw = dict(rhs="baseline", t_span=(0.0, 10.0), y0=[20.0, 20.0], method="RK23", rtol=1e-7, atol=1e-8)
    r = oracle.solve(max_num_steps=5, **kw)
    assert not r.success and r.status == -52
    r = oracle.solve(max_ram_MB=0, **kw)
    assert not r.success and r.status == -53


def test_survey_known_answers(oracle):
    """values read off the compiled reference at survey time."""
    r = oracle.solve("lotkavolterra", (0, 10), [10, 5], "rk45", 1e-7, 1e-9, args=[1.5, 1, 3, 1])
    assert (r.status, r.size, r.steps_taken) == (1, 214, 213)
    assert r.t[1] == 0.00883576448311094
    assert r.y[:, -1].tolist() == [0.2872127983010663, 0.44977767381276684]
    r = oracle.solve("lorenz", (0, 10), [1, 0, 0], "dop853", 1e-9, 1e-11, args=[10, 28, 8 / 3])
    assert (r.size, r.steps_taken) == (251, 250) and r.t[1] == 0.00035360457856101776
    assert r.y[:, -1].tolist() == [-5.857685408025666, -5.831082511317166, 23.932133016385468]
    r = oracle.solve("lorenz", (0, 10), [1, 0, 0], "dop853", 1e-9, 1e-11, args=[10, 28, 8 / 3],
                     t_eval=np.linspace(0, 10, 1000), dense_output=True, call_t=4.335)
    assert (r.size, r.steps_taken) == (1000, 250)
    assert r.y[:, 500].tolist() == [-6.978473226436477, -7.0878798894776285, 25.03125408292453]
    assert r.dense.ravel().tolist() == [-7.164745326341834, -6.681525540917222, 26.09281033210707]
    r = oracle.solve("pendulum", (0, 10), [0.01, 0], "rk23", 1e-6, 1e-8, args=[1, 1, 9.81],
                     events=["y0_gt_50", "y1_lt_50", "t_gt_5"])
    assert (r.size, r.steps_taken) == (1063, 1062)
    assert [len(t) for t in r.t_events] == [0, 1062, 529] and r.t_events[2][0] == 5.0066788312323816
    r = oracle.solve("pendulum_drag", (0, 10), [1, 0], "rk23", 1e-6, 1e-8, args=[1, 1, 9.81, 1.5, 6.5],
                     events=[("y0", 0, None), ("y1", -1, None), ("t_minus_arg4", 0, 1)])
    assert (r.status, r.size, r.steps_taken, r.t[-1], r.event_terminate_index) == (2, 851, 850, 6.5, 2)
    assert [len(t) for t in r.t_events] == [8, 4, 1]
    assert r.t_events[0][:2].tolist() == [0.5051472489393665, 1.3347286902118365]
    assert r.t_events[1][:2].tolist() == [0.0, 1.7166667158465867]
    assert np.allclose(r.y_events[2].ravel(), [0.0060083463438988117, 0.044020076765382436, 0.021511998808781553], rtol=0, atol=1e-15)


def test_quirk11_fix_only_touches_last_row(oracle):
    """with fix_quirk11 the final row equals the terminal event's y_events entry."""
    kw = dict(rhs="pendulum_drag", t_span=(0, 10), y0=[1, 0], method="rk23", rtol=1e-6, atol=1e-8, args=[1, 1, 9.81, 1.5, 6.5],
              events=[("y0", 0, None), ("y1", -1, None), ("t_minus_arg4", 0, 1)])
    a = oracle.solve(**kw)
    b = oracle.solve(fix_quirk11=True, **kw)
    assert np.array_equal(a.t, b.t) and np.array_equal(a.y[:, :-1], b.y[:, :-1])
    assert not np.array_equal(a.y[:, -1], b.y[:, -1])
    assert np.array_equal(b.y[:, -1], b.y_events[2][:, 0])
