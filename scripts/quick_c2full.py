"""C2full (Lorenz DOP853, 1000 t_eval samples + dense output) device-resident kernel time at QB trajectories."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c, bench
for name in ("C2full", "C2"):
    w = bench.WORKLOADS[name]; B = int(os.environ.get("QB", 1 << 20))
    y0, args = bench.make_inputs(w, B, 100)
    y0d, ad = torch.from_numpy(y0).cuda(), torch.from_numpy(args).cuda()
    kw = bench.solve_kwargs(c, w, 0)
    res = None; best = 1e9
    for _ in range(4):
        res = c.cysolve_ivp_batch(w["rhs"], (0.0, 10.0), y0d, args=ad, solution_reuse=res, **kw)
        best = min(best, res.kernel_ms)
    print(f"{name} B={B} kernel_ms={best:.3f} launches={res.launches}", flush=True)
    res.close()
