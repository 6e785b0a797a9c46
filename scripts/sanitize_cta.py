"""Small exercise of the CTA-per-trajectory kernels for compute-sanitizer (memcheck / racecheck / synccheck):
register-resident kernel at C = 1, 2, 4, 8 with ragged component counts, both FP builds, t_eval incl. DOP853's
extra stages, the shared-memory/slab kernel, and a user right-hand side through NVRTC."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cyrk_b200 as c
rng = np.random.default_rng(0)
for n in (130, 700, 1300, 2500):
    yl = 100.0 * (1 + 0.01 * rng.uniform(-1, 1, (3, n)))
    for m in ("RK23", "RK45", "DOP853"):
        for strict in (False, True):
            r = c.cysolve_ivp_batch("largey", (0, 1), yl, method=m, args=[-0.5, float(n)], rtol=1e-5, atol=1e-8, expected_size=2, strict_fp=strict)
            assert (r.status == 1).all()
        r = c.cysolve_ivp_batch("largey", (0, 1), yl, method=m, args=[-0.5, float(n)], rtol=1e-5, atol=1e-8, t_eval=np.linspace(0, 1, 4))
        r = c.cysolve_ivp_batch("largey_simple", (1, 0), yl, method=m, args=[-0.5, float(n)], rtol=1e-5, atol=1e-8, store="final")
c.lib().cyrk_b200_set_option(b"cta_kernel", 1)
yl = 100.0 * (1 + 0.01 * rng.uniform(-1, 1, (3, 300)))
r = c.cysolve_ivp_batch("largey", (0, 1), yl, method="RK45", args=[-0.5, 300.0], rtol=1e-5, atol=1e-8, t_eval=np.linspace(0, 1, 4))
c.lib().cyrk_b200_set_option(b"cta_kernel", 0)
src = r'''
__device__ double chain(unsigned i, double t, const double* y, unsigned n, const char* args) {
    const double r = ((const double*)args)[0];
    const double left = i > 0 ? y[i - 1] : 0.0, right = i + 1 < n ? y[i + 1] : 0.0;
    return r * (left - 2.0 * y[i] + right) + 0.1 * sin(t);
}
'''
u = c.UserRHS(src, "chain", n_y=0, args_bytes=8)
y0 = rng.uniform(0, 1, (3, 700))
r = c.cysolve_ivp_batch(u, (0, 1), y0, method="RK45", args=np.array([[0.5], [0.8], [1.1]]), rtol=1e-5, atol=1e-8)
assert (r.status == 1).all()
print("sanitize cta run ok")
