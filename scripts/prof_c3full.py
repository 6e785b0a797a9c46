"""One C3full solve (damped pendulum RK23, 1 extra, 3 events, store=final, 256 Ki) for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
rng = np.random.default_rng(2)
B = int(os.environ.get("QB", 1 << 18))
y0 = torch.from_numpy(np.stack([rng.uniform(-1.5, 1.5, B), rng.uniform(-1, 1, B)], 1)).cuda()
args = torch.from_numpy(np.stack([np.ones(B), np.ones(B), np.full(B, 9.81), np.full(B, 1.5), rng.uniform(2.0, 12.0, B)], 1)).cuda()
ts = torch.tensor([0.0, 10.0], dtype=torch.float64, device="cuda")
ev = [c.Event("y0"), c.Event("y1", direction=-1), c.Event("t_minus_arg4", terminal=True)]
res = c.cysolve_ivp_batch("pendulum_drag", ts, y0, method="RK23", args=args, rtol=1e-6, atol=1e-8, num_extra=1, store="final", events=ev)
print("kernel_ms", res.kernel_ms)
