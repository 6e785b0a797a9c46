"""Same Lotka-Volterra RK45 batch through the precompiled built-in and through an NVRTC user right-hand side (device-resident)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
rng = np.random.default_rng(0)
B = int(os.environ.get("QB", 1 << 22))
y0 = torch.from_numpy(np.stack([rng.uniform(5, 15, B), rng.uniform(2.5, 7.5, B)], 1)).cuda()
ts = torch.tensor([0., 10.], dtype=torch.float64, device="cuda")
ad = torch.tensor([1.5, 1, 3, 1], dtype=torch.float64, device="cuda")
SRC_ARGS = "__device__ void my_lv(double* dy, double t, const double* y, const char* args) { const double* a = (const double*)args; dy[0] = a[0] * y[0] - a[1] * y[0] * y[1]; dy[1] = -a[2] * y[1] + a[3] * y[0] * y[1]; }"
SRC_LIT = "__device__ void lv2(double* dy, double t, const double* y, const char* a) { dy[0] = y[0] * (1.5 - 1.0 * y[1]); dy[1] = y[1] * (1.0 * y[0] - 3.0); }"
cases = [("builtin", "lotkavolterra", dict(args=ad)),
         ("user+args", c.UserRHS(SRC_ARGS, "my_lv", n_y=2, args_bytes=32), dict(args=ad)),
         ("user literal", c.UserRHS(SRC_LIT, "lv2", n_y=2), {})]
for name, rhs, kw in cases:
    best = 1e9
    for rep in range(4):
        r = c.cysolve_ivp_batch(rhs, ts, y0, method="RK45", rtol=1e-7, atol=1e-9, store="final", **kw)
        best = min(best, r.kernel_ms)
    print(f"{name:14s} B={B} kernel_ms={best:.3f} total_ms={r.total_ms:.3f} launches={r.launches}", flush=True)
