"""Small end-to-end exercise of every kernel family for compute-sanitizer memcheck."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cyrk_b200 as c
rng = np.random.default_rng(0)
B = 257
y0 = np.stack([rng.uniform(5, 15, B), rng.uniform(2.5, 7.5, B)], 1)
for m in ("RK23", "RK45", "DOP853"):
    for strict in (False, True):
        r = c.cysolve_ivp_batch("lotkavolterra", (0, 3), y0, method=m, args=[1.5, 1, 3, 1], rtol=1e-5, atol=1e-8, strict_fp=strict, store="rows", expected_size=8)
        r = c.cysolve_ivp_batch("lotkavolterra", (0, 3), y0, method=m, args=[1.5, 1, 3, 1], rtol=1e-5, atol=1e-8, strict_fp=strict, store="final")
        r = c.cysolve_ivp_batch("lotkavolterra", (3, 0), y0, method=m, args=[1.5, 1, 3, 1], rtol=1e-5, atol=1e-8, t_eval=np.linspace(3, 0, 7), dense_output=True)
        out, codes = r.call(np.linspace(0, 3, 5))
yp = np.stack([rng.uniform(-1.5, 1.5, B), rng.uniform(-1, 1, B)], 1)
args = np.stack([np.ones(B), np.ones(B), np.full(B, 9.81), np.full(B, 1.5), rng.uniform(0.5, 4.0, B)], 1)
ev = [c.Event("y0"), c.Event("y1", direction=-1), c.Event("t_minus_arg4", terminal=True), c.Event("t_gt_5")]
for m in ("RK23", "DOP853"):
    r = c.cysolve_ivp_batch("pendulum_drag", (0, 3), yp, method=m, args=args, rtol=1e-5, atol=1e-8, events=ev, num_extra=1, expected_size=4, dense_output=True)
    r = c.cysolve_ivp_batch("pendulum_drag", (0, 3), yp, method=m, args=args, rtol=1e-5, atol=1e-8, events=ev, num_extra=1, t_eval=np.linspace(0, 3, 11))
n = 130
yl = 100.0 * (1 + 0.01 * rng.uniform(-1, 1, (5, n)))
for m in ("RK23", "RK45", "DOP853"):
    r = c.cysolve_ivp_batch("largey", (0, 2), yl, method=m, args=[-0.5, float(n)], rtol=1e-5, atol=1e-8, expected_size=2)
    r = c.cysolve_ivp_batch("largey", (0, 2), yl, method=m, args=[-0.5, float(n)], rtol=1e-5, atol=1e-8, t_eval=np.linspace(0, 2, 5))
src = "__device__ void f(double* dy, double t, const double* y, const char* a) { dy[0] = -y[0] + sin(t); dy[1] = y[0]; dy[2] = y[1] * 2.0; }"
u = c.UserRHS(src, "f", n_y=2, n_extra=1)
r = c.cysolve_ivp_batch(u, (0, 2), y0, method="RK45", num_extra=1, rtol=1e-5, atol=1e-8, t_eval=np.linspace(0, 2, 5))
# step-parallel interpolants: precompiled (Lorenz: 3 components, padded step records) and NVRTC, forward t_eval beyond both ends
yz = rng.uniform(1.0, 8.0, (B, 3))
for m in ("RK23", "RK45", "DOP853"):
    r = c.cysolve_ivp_batch("lorenz", (0, 2), yz, method=m, args=[10.0, 28.0, 8 / 3], rtol=1e-6, atol=1e-9, t_eval=np.linspace(-0.1, 2.1, 97), dense_output=True, expected_size=8)
    r = c.cysolve_ivp_batch("lorenz", (0, 2), yz, method=m, args=[10.0, 28.0, 8 / 3], rtol=1e-6, atol=1e-9, dense_output=True)
c.lib().cyrk_b200_set_option(b"two_pass", 2)
r = c.cysolve_ivp_batch("lorenz", (0, 2), yz, method="DOP853", args=[10.0, 28.0, 8 / 3], rtol=1e-6, atol=1e-9, t_eval=np.linspace(0, 2, 33))
c.lib().cyrk_b200_set_option(b"two_pass", 1)
u2 = c.UserRHS("__device__ void g(double* dy, double t, const double* y, const char* a) { dy[0] = -y[0] + sin(t); dy[1] = y[0]; }", "g", n_y=2)
r = c.cysolve_ivp_batch(u2, (0, 2), y0, method="DOP853", rtol=1e-6, atol=1e-9, t_eval=np.linspace(0, 2, 21), dense_output=True)
print("sanitize run ok")
