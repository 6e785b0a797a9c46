"""C3full (events, store='final') device-resident kernel time: one-pass (CYRK_EV2=0) vs two-pass events (CYRK_EV2=1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
import bench
ev2 = int(os.environ.get("CYRK_EV2", "1"))
c.lib().cyrk_b200_set_option(b"events_two_pass", ev2)
w = bench.WORKLOADS["C3full"]
B = int(os.environ.get("QB", 1 << 20))
y0, args = bench.make_inputs(w, B, 100)
y0d, ad = torch.from_numpy(y0).cuda(), torch.from_numpy(args).cuda()
kw = bench.solve_kwargs(c, w, 0)
res = None; best = 1e9; tot = 1e9
for _ in range(5):
    res = c.cysolve_ivp_batch(w["rhs"], (0.0, 10.0), y0d, args=ad, solution_reuse=res, **kw)
    best = min(best, res.kernel_ms); tot = min(tot, res.total_ms)
print(f"C3full B={B} events_two_pass={ev2}: kernel_ms={best:.2f} total_ms={tot:.2f} launches={res.launches} -> {B / tot * 1e3:.3e} solves/s", flush=True)
