"""Events + storage (VAR_FULL) on the pendulum: RK23 / RK45, store='rows'; A/B for the register cap (CYRK_B200_LIB)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
rng = np.random.default_rng(2)
B = 1 << 18
y0 = torch.from_numpy(np.stack([rng.uniform(-1.5, 1.5, B), rng.uniform(-1, 1, B)], 1)).cuda()
args = torch.from_numpy(np.stack([np.ones(B), np.ones(B), np.full(B, 9.81), np.full(B, 1.5), rng.uniform(2.0, 12.0, B)], 1)).cuda()
ts = torch.tensor([0.0, 10.0], dtype=torch.float64, device="cuda")
ev = [c.Event("y0"), c.Event("y1", direction=-1), c.Event("t_minus_arg4", terminal=True)]
for method, rtol in (("RK23", 1e-6), ("RK45", 1e-8), ("DOP853", 1e-10)):
    res = None; best = 1e9
    for rep in range(3):
        res = c.cysolve_ivp_batch("pendulum_drag", ts, y0, method=method, args=args, rtol=rtol, atol=rtol * 1e-2, num_extra=1,
                                  store="rows", t_eval=np.linspace(0, 10, 21), events=ev, solution_reuse=res)
        best = min(best, res.kernel_ms)
    print(f"{method} FULL (events + t_eval rows) kernel_ms={best:7.2f}", flush=True)
    res.close()
