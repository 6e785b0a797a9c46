"""C3full (damped pendulum RK23, 1 extra, 3 events, store=final, 1 Mi): VAR_FULL vs VAR_EVENTS kernels, root_batch sweep."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
rng = np.random.default_rng(2)
B = 1 << 20
y0 = torch.from_numpy(np.stack([rng.uniform(-1.5, 1.5, B), rng.uniform(-1, 1, B)], 1)).cuda()
args = torch.from_numpy(np.stack([np.ones(B), np.ones(B), np.full(B, 9.81), np.full(B, 1.5), rng.uniform(2.0, 12.0, B)], 1)).cuda()
ts = torch.tensor([0.0, 10.0], dtype=torch.float64, device="cuda")
ev = [c.Event("y0"), c.Event("y1", direction=-1), c.Event("t_minus_arg4", terminal=True)]
for variant in (1,):
    c.lib().cyrk_b200_set_option(b"events_variant", variant)
    for rb in (1, 2, 4, 6, 8, 10, 12):
        c.lib().cyrk_b200_set_option(b"root_batch", rb)
        res = None; best = 1e9
        for rep in range(3):
            res = c.cysolve_ivp_batch("pendulum_drag", ts, y0, method="RK23", args=args, rtol=1e-6, atol=1e-8, num_extra=1,
                                      store="final", events=ev, solution_reuse=res)
            best = min(best, res.kernel_ms)
        print(f"events_variant={variant} root_batch={rb:2d} kernel_ms={best:7.2f} solves/s={B / best * 1e3:.3e}", flush=True)
        res.close()
c.lib().cyrk_b200_set_option(b"events_variant", 1); c.lib().cyrk_b200_set_option(b"root_batch", 12)
