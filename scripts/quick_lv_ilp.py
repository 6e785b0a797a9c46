"""Hypothesis test for 'two trajectories per thread': an NVRTC system made of TWO independent Lotka-Volterra pairs (n_y = 4,
one shared step size -- not the same numerics, only the same instruction mix at twice the ILP) against the n_y = 2 user
system, both through the NVRTC path; prints ms and FP64-equivalent throughput per attempt."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
rng = np.random.default_rng(0)
B = int(os.environ.get("QB", 1 << 22))
src2 = "__device__ void lv2(double* dy, double t, const double* y, const char* a) { dy[0] = y[0] * (1.5 - 1.0 * y[1]); dy[1] = y[1] * (1.0 * y[0] - 3.0); }"
src4 = ("__device__ void lv4(double* dy, double t, const double* y, const char* a) { dy[0] = y[0] * (1.5 - 1.0 * y[1]); dy[1] = y[1] * (1.0 * y[0] - 3.0);"
        " dy[2] = y[2] * (1.5 - 1.0 * y[3]); dy[3] = y[3] * (1.0 * y[2] - 3.0); }")
ts = torch.tensor([0., 10.], dtype=torch.float64, device="cuda")
for name, src, ny, flop in (("lv2", src2, 2, 2 * 62 + 6 * 8 + 22), ("lv4", src4, 4, 4 * 62 + 6 * 16 + 22)):
    u = c.UserRHS(src, name, n_y=ny)
    y0 = np.stack([rng.uniform(5, 15, B), rng.uniform(2.5, 7.5, B)] * (ny // 2), 1)
    y0d = torch.from_numpy(np.ascontiguousarray(y0)).cuda()
    best = 1e9
    for rep in range(3):
        r = c.cysolve_ivp_batch(u, ts, y0d, method="RK45", rtol=1e-7, atol=1e-9, store="final")
        best = min(best, r.kernel_ms)
    att = int(torch.as_tensor(r.steps_taken, device="cuda").sum().item()) + int(torch.as_tensor(r.n_rejected, device="cuda").sum().item())
    print(f"{name}: B={B} kernel_ms={best:.3f} attempts={att} ns/attempt/SM-lane... attempts/s={att / best * 1e3:.3e} alg_TFLOPs={att * flop / best * 1e-9:.2f}", flush=True)
