"""Golden-vector cases shared by make_golden.py (generator, runs the compiled reference) and the tests.

Each case is a dict of keyword arguments of oracle.solve() (== the cysolve_ivp argument surface).
Kept small: the whole fixture is a few hundred kB.
"""
import struct

import numpy as np

LV = dict(rhs="lotkavolterra", y0=[10.0, 5.0], args=[1.5, 1.0, 3.0, 1.0])
LORENZ = dict(rhs="lorenz", y0=[1.0, 0.0, 0.0], args=[10.0, 28.0, 8.0 / 3.0])
LORENZ_X = dict(rhs="lorenz_extraoutput", y0=[1.0, 0.0, 0.0], args=[10.0, 28.0, 8.0 / 3.0])
PEND = dict(rhs="pendulum", y0=[0.01, 0.0], args=[1.0, 1.0, 9.81])
DRAG = dict(rhs="pendulum_drag", y0=[1.0, 0.0], args=[1.0, 1.0, 9.81, 1.5, 6.5])
EXTRA = dict(rhs="extraoutput", y0=[20.0, 20.0])   # periodic, so it can be integrated backward too
ARB_FAIL = struct.pack("d??6xdd", 1.0, True, True, 1.0, 9.81)
ARB_OK = struct.pack("d??6xdd", 1.0, False, True, 1.0, 9.81)


def cases():
    c = {}
    # SURVEY.md 8c known answers (C1/C2/C3 style)
    c["lv_rk45_c1"] = dict(LV, t_span=(0, 10), method="rk45", rtol=1e-7, atol=1e-9)
    c["lorenz_dop853_c2"] = dict(LORENZ, t_span=(0, 10), method="dop853", rtol=1e-9, atol=1e-11)
    c["lorenz_dop853_c2_teval_dense"] = dict(LORENZ, t_span=(0, 10), method="dop853", rtol=1e-9, atol=1e-11,
                                             t_eval=np.linspace(0, 10, 101), dense_output=True,
                                             call_t=[0.0, 0.3, 4.335, 9.99, 10.0])
    c["pend_rk23_c3"] = dict(PEND, t_span=(0, 10), method="rk23", rtol=1e-6, atol=1e-8)
    c["pend_rk23_builtin_events"] = dict(PEND, t_span=(0, 10), method="rk23", rtol=1e-6, atol=1e-8,
                                         events=["y0_gt_50", "y1_lt_50", "t_gt_5"])
    c["drag_rk23_c3_events"] = dict(DRAG, t_span=(0, 10), method="rk23", rtol=1e-6, atol=1e-8,
                                    events=[("y0", 0, None), ("y1", -1, None), ("t_minus_arg4", 0, 1)])
    c["drag_rk23_c3_events_teval"] = dict(DRAG, t_span=(0, 10), method="rk23", rtol=1e-6, atol=1e-8,
                                          t_eval=np.linspace(0, 10, 21),
                                          events=[("y0", 0, None), ("y1", -1, None), ("t_minus_arg4", 0, 1)])
    # every method x direction x (t_eval, dense) on two problems
    for m in ("rk23", "rk45", "dop853"):
        for span in ((0, 10), (10, 0)):
            tag = "fwd" if span[1] > span[0] else "bwd"
            c[f"lv_{m}_{tag}"] = dict(LV, t_span=span, method=m, rtol=1e-6, atol=1e-8)
            c[f"lv_{m}_{tag}_teval"] = dict(LV, t_span=span, method=m, rtol=1e-6, atol=1e-8,
                                            t_eval=np.linspace(span[0], span[1], 17))
            c[f"extra_{m}_{tag}_dense"] = dict(EXTRA, t_span=span, method=m, rtol=1e-6, atol=1e-8, dense_output=True,
                                                 call_t=np.linspace(min(span), max(span), 9))
            c[f"extra_{m}_{tag}_teval_dense"] = dict(EXTRA, t_span=span, method=m, rtol=1e-6, atol=1e-8,
                                                       t_eval=np.linspace(span[0], span[1], 13), dense_output=True,
                                                       call_t=np.linspace(min(span), max(span), 9))
        # events_test.pyx-style Lorenz events, with and without termination
        for term in (False, True):
            evs = [("lorenz_t_ge_5", 0, 1 if term else None), ("lorenz_y_range", 0, None), ("lorenz_args", 0, None)]
            c[f"lorenz_{m}_events_term{int(term)}"] = dict(LORENZ, t_span=(0, 10), method=m, rtol=1e-6, atol=1e-6, events=evs)
            c[f"lorenzx_{m}_events_term{int(term)}_teval"] = dict(LORENZ_X, t_span=(0, 10), method=m, rtol=1e-6, atol=1e-6,
                                                                  events=evs, t_eval=np.linspace(0, 10, 11))
        # knobs and status codes
        c[f"lv_{m}_first_step"] = dict(LV, t_span=(0, 10), method=m, rtol=1e-6, atol=1e-8, first_step=1e-3)
        c[f"lv_{m}_max_step"] = dict(LV, t_span=(0, 10), method=m, rtol=1e-6, atol=1e-8, max_step=0.5)
        c[f"lv_{m}_max_num_steps"] = dict(LV, t_span=(0, 10), method=m, rtol=1e-6, atol=1e-8, max_num_steps=5)
        c[f"lv_{m}_tiny_ram"] = dict(LV, t_span=(0, 1000), method=m, rtol=1e-6, atol=1e-8, max_ram_MB=0)
        c[f"lv_{m}_array_tols"] = dict(LV, t_span=(0, 10), method=m, rtol=[1e-6, 1e-9], atol=[1e-8, 1e-10])
        c[f"arb_{m}_fail54"] = dict(rhs="arbitrary_args", y0=[0.01, 0.0], args=ARB_FAIL, t_span=(0, 10), method=m,
                                    rtol=1e-6, atol=1e-8)
        c[f"arb_{m}_ok"] = dict(rhs="arbitrary_args", y0=[0.01, 0.0], args=ARB_OK, t_span=(0, 10), method=m,
                                rtol=1e-6, atol=1e-8)
        c[f"preeval_{m}"] = dict(rhs="pendulum_preeval", y0=[0.01, 0.0], args=[1.0, 1.0, 9.81], t_span=(0, 10), method=m,
                                 rtol=1e-6, atol=1e-8)
        c[f"extra_{m}_teval_short"] = dict(rhs="extraoutput", y0=[20.0, 20.0], t_span=(0, 10), method=m, rtol=1e-5, atol=1e-8,
                                           t_eval=np.linspace(0, 5, 7))
        c[f"accuracy_{m}"] = dict(rhs="accuracy", y0=[0.0, 1.0], t_span=(0, 10), method=m, rtol=1e-8, atol=1e-9,
                                  dense_output=True, call_t=np.linspace(0.1, 0.4, 10))
    # the reference's own dense + extra known-answer test (Tests/D_PySolver_Tests/test_b_pysolve_dense.py)
    c["dense_kat"] = dict(rhs="dense_kat", y0=[5.0, 2.0], t_span=(0, 10), method="rk45", rtol=1e-3, atol=1e-6,
                          max_step=100000.0, dense_output=True, call_t=[0.3, 4.0, 8.0, 9.10, 9.9])
    # small large-n_y systems for the CTA-per-trajectory path
    c["largey_rk45_n96"] = dict(rhs="largey", y0=100.0 * np.ones(96), args=[-0.5, 96.0], t_span=(0, 10), method="rk45",
                                rtol=1e-6, atol=1e-8)
    c["largey_simple_rk45_n64"] = dict(rhs="largey_simple", y0=100.0 * np.ones(64), args=[-0.5, 64.0], t_span=(0, 10),
                                       method="rk45", rtol=1e-6, atol=1e-8)
    return c
