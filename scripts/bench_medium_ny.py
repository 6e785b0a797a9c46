"""Thread-per-trajectory path at 4 <= n_y <= 32 (NVRTC spring chains, tests/medium_ny.py): solves/s, algorithmic FP64 fraction."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import cyrk_b200 as c
import bench
from medium_ny import chain_source, chain_inputs
peak = c.lib().cyrk_b200_measure_fp64_tflops(0)
print(f"fp64 probe {peak:.2f} TFLOP/s")
for ny in (4, 8, 16, 32):
    m = ny // 2
    u = c.UserRHS(chain_source(m), f"chain{ny}", n_y=ny, args_bytes=16)
    B = (1 << 20) // (ny // 4)
    y0, args = chain_inputs(m, B, 7)
    y0d, ad = torch.from_numpy(y0).cuda(), torch.from_numpy(args).cuda()
    for method in ("RK23", "RK45", "DOP853"):
        rtol, atol = (1e-5, 1e-8) if method == "RK23" else (1e-8, 1e-10)
        best = 1e9
        for _ in range(3):
            r = c.cysolve_ivp_batch(u, (0.0, 6.0), y0d, method=method, args=ad, rtol=rtol, atol=atol, store="final")
            best = min(best, r.kernel_ms)
        att = int(torch.as_tensor(r.steps_taken, device="cuda").sum().item()) + int(torch.as_tensor(r.n_rejected, device="cuda").sum().item())
        f_rhs = 6 * m       # per evaluation: m x (2 mul + 2 sub + mul + sub)
        f_att = ny * bench.C_M[method] + bench.E_M[method] * f_rhs + bench.S_CTRL[method]
        tf = att * f_att / (best * 1e-3) / 1e12
        print(f"n_y={ny:2d} {method:6s} B={B} kernel_ms={best:.2f} solves/s={B / best * 1e3:.3e} attempts={att} alg TFLOP/s={tf:.2f} ({tf / peak:.3f} of FP64 peak)", flush=True)

# DOP853 with 16 < n_y <= 32: warp per trajectory (cta_core.cuh, option small_cta) vs thread per trajectory (default)
for ny in (24, 32):
    m = ny // 2
    u = c.UserRHS(chain_source(m), f"chain{ny}", n_y=ny, args_bytes=16)
    B = (1 << 20) // (ny // 4)
    y0, args = chain_inputs(m, B, 7)
    y0d, ad = torch.from_numpy(y0).cuda(), torch.from_numpy(args).cuda()
    for small in (1, 0):
        c.lib().cyrk_b200_set_option(b"small_cta", small)
        best = 1e9
        for _ in range(3):
            r = c.cysolve_ivp_batch(u, (0.0, 6.0), y0d, method="DOP853", args=ad, rtol=1e-8, atol=1e-10, store="final")
            best = min(best, r.kernel_ms)
        att = int(torch.as_tensor(r.steps_taken, device="cuda").sum().item()) + int(torch.as_tensor(r.n_rejected, device="cuda").sum().item())
        f_att = ny * bench.C_M["DOP853"] + bench.E_M["DOP853"] * 6 * m + bench.S_CTRL["DOP853"]
        tf = att * f_att / (best * 1e-3) / 1e12
        print(f"n_y={ny:2d} DOP853 small_cta={small} B={B} kernel_ms={best:.2f} solves/s={B / best * 1e3:.3e} attempts={att} alg TFLOP/s={tf:.2f} ({tf / peak:.3f} of FP64 peak)", flush=True)
c.lib().cyrk_b200_set_option(b"small_cta", 0)
