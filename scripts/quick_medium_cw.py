"""Spring chain with n_y = 2m components, DOP853: whole-vector form on the thread-per-trajectory path vs the component-wise
form on the register-resident CTA kernel (one thread per component, several CTAs per SM)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import cyrk_b200 as c
from medium_ny import chain_source, chain_inputs
def cw_source(m):
    return f"""
__device__ double chaincw{2*m}(unsigned i, double t, const double* y, unsigned n, const char* args) {{
    const double* a = (const double*)args;
    const double k = a[0], c = a[1];
    if (i < {m}u) return y[{m} + i];
    const unsigned j = i - {m}u;
    const double left = j > 0 ? y[j - 1] : 0.0, right = j + 1 < {m}u ? y[j + 1] : 0.0;
    return -k * (2.0 * y[j] - left - right) - c * y[{m} + j];
}}
"""
for ny in (16, 32, 64, 128):
    m = ny // 2
    B = (1 << 20) // (ny // 4)
    y0, args = chain_inputs(m, B, 7)
    y0d, ad = torch.from_numpy(y0).cuda(), torch.from_numpy(args).cuda()
    uw = c.UserRHS(chain_source(m), f"chain{ny}", n_y=ny, args_bytes=16)
    uc = c.UserRHS(cw_source(m), f"chaincw{ny}", n_y=0, args_bytes=16)
    for method in ("RK45", "DOP853"):
        ref = None
        for tag, u, per_sm in (("whole-vector", uw, 0), ("component-wise per_sm=8", uc, 8), ("component-wise per_sm=16", uc, 16), ("component-wise per_sm=32", uc, 32)):
            c.lib().cyrk_b200_set_option(b"cta2_per_sm", per_sm)
            best = 1e9
            for _ in range(3):
                r = c.cysolve_ivp_batch(u, (0.0, 6.0), y0d, method=method, args=ad, rtol=1e-8, atol=1e-10, store="final")
                best = min(best, r.kernel_ms)
            st = torch.as_tensor(r.steps_taken, device="cuda").cpu().numpy()
            yl = torch.as_tensor(r.y_last, device="cuda").cpu().numpy()
            d = ""
            if ref is None: ref = (st.copy(), yl.copy())
            else: d = f" steps_equal={np.mean(st == ref[0]):.4f} max|dy|={np.abs(yl - ref[1]).max():.2e}"
            print(f"n_y={ny:3d} {method:6s} {tag:26s} B={B} kernel_ms={best:.2f} solves/s={B / best * 1e3:.3e}{d}", flush=True)
c.lib().cyrk_b200_set_option(b"cta2_per_sm", 0)
