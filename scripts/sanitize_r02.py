"""Small end-to-end exercise of the round-2 kernels and host paths for compute-sanitizer memcheck: two-pass events (generic and
event-list-specialised detection solve, queue overflow re-run), packed event lists, streamed solves, step records as the
dense store, per-trajectory t_eval / tolerances, extra outputs on the CTA path, device-side counter maxima."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import cyrk_b200 as c
L = c.lib()
rng = np.random.default_rng(0)
B = 1100
yp = np.stack([rng.uniform(-1.5, 1.5, B), rng.uniform(-1, 1, B)], 1)
args = np.stack([np.ones(B), np.ones(B), np.full(B, 9.81), np.full(B, 1.5), rng.uniform(0.5, 4.0, B)], 1)
ev = [c.Event("y0"), c.Event("y1", direction=-1), c.Event("t_minus_arg4", terminal=True), c.Event("t_gt_5")]
for static in (0, 1):
    L.cyrk_b200_set_option(b"events_static", static)
    for m in ("RK23", "DOP853"):
        r = c.cysolve_ivp_batch("pendulum_drag", (0, 6), yp, method=m, args=args, rtol=1e-5, atol=1e-8, events=ev, num_extra=1, store="final")
        assert r.event_total > 0 and len(r[3].t_events) == 4
# queue overflow (a step-function event triggers on every step)
r = c.cysolve_ivp_batch("pendulum", (0, 10), np.tile([0.3, 0.0], (40, 1)), method="RK45", args=[1.0, 1.0, 9.81], rtol=1e-8, atol=1e-10,
                        events=[c.Event("t_gt_5"), c.Event("y0")], store="final", expected_size=1500)
# streamed: rows + events, t_eval + dense records
L.cyrk_b200_set_option(b"stream_chunk", 256)
r = c.cysolve_ivp_batch("pendulum_drag", (0, 6), yp, method="RK23", args=args, rtol=1e-5, atol=1e-8, events=ev, num_extra=1, store="rows", expected_size=600)
yz = rng.uniform(1.0, 8.0, (B, 3))
r = c.cysolve_ivp_batch("lorenz", (0, 2), yz, method="DOP853", args=[10.0, 28.0, 8 / 3], rtol=1e-6, atol=1e-9, t_eval=np.linspace(-0.1, 2.0, 97), dense_output=True, expected_size=8)
out, codes = r.call(np.linspace(0, 2, 9))
r = c.cysolve_ivp_batch("lorenz", (0, 2), yz, method="RK45", args=[10.0, 28.0, 8 / 3], rtol=1e-6, atol=1e-9, dense_output=True, store="rows", expected_size=8)
out, codes = r.call(np.linspace(0, 2, 9))
L.cyrk_b200_set_option(b"stream_chunk", 0)
# per-trajectory tolerances (NVRTC feature build of a built-in) and t_eval, through the pipelined path as well
y0 = np.stack([rng.uniform(5, 15, B), rng.uniform(2.5, 7.5, B)], 1)
rt = 10.0 ** rng.uniform(-8, -5, (B, 2)); at = rt * 1e-2
te = np.sort(rng.uniform(0, 3, (B, 11)), axis=1)
r = c.cysolve_ivp_batch("lotkavolterra", (0, 3), y0, method="RK45", args=[1.5, 1, 3, 1], rtol=rt, atol=at, t_eval=te)
L.cyrk_b200_set_option(b"pipe_min", 256); L.cyrk_b200_set_option(b"pipe_chunk", 1024)
r = c.cysolve_ivp_batch("lotkavolterra", (0, 3), y0, method="DOP853", args=[1.5, 1, 3, 1], rtol=rt, atol=at, store="final")
_ = np.array(r.n_rejected), np.array(r.t_last)
L.cyrk_b200_set_option(b"pipe_min", 1 << 18); L.cyrk_b200_set_option(b"pipe_chunk", 1 << 21)
# extra outputs on the CTA path (register-resident and slab kernels)
from test_user_rhs import CHAIN_X_SRC
u = c.UserRHS(CHAIN_X_SRC, "chainx", n_y=0, n_extra=2, args_bytes=8)
for n in (130, 4200):
    yl = rng.uniform(0.0, 1.0, (3, n))
    for m in ("RK23", "RK45", "DOP853"):
        r = c.cysolve_ivp_batch(u, (0, 1), yl, method=m, args=[[0.5], [0.8], [1.1]], rtol=1e-5, atol=1e-8, num_extra=2, store="rows", expected_size=4)
print("sanitize run ok")
