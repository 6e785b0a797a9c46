"""largey n_y = 4096 with all three methods on the CTA path (register-resident vs shared-memory/slab kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
rng = np.random.default_rng(3)
n, B = 4096, 1184
y0d = torch.from_numpy(100.0 * (1 + 0.01 * rng.uniform(-1, 1, (B, n)))).cuda()
ts = torch.tensor([0.0, 10.0], dtype=torch.float64, device="cuda")
ar = torch.tensor([-0.5, float(n)], dtype=torch.float64, device="cuda")
for method in ("RK23", "RK45", "DOP853"):
    for kernel in (2, 1):
        c.lib().cyrk_b200_set_option(b"cta_kernel", kernel)
        res = None; best = 1e9
        for rep in range(3):
            res = c.cysolve_ivp_batch("largey", ts, y0d, method=method, args=ar, rtol=1e-6, atol=1e-8, store="final", solution_reuse=res)
            best = min(best, res.kernel_ms)
        st = torch.as_tensor(res.steps_taken, device="cuda").float().mean().item()
        nr = torch.as_tensor(res.n_rejected, device="cuda").float().mean().item()
        print(f"{method:7s} kernel={'cta2' if kernel == 2 else 'cta '} kernel_ms={best:.2f} steps={st:.1f} rejected={nr:.1f}", flush=True)
        res.close()
c.lib().cyrk_b200_set_option(b"cta_kernel", 0)
