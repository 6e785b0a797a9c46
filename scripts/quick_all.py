"""Kernel/total times of the main workloads for A/B comparisons of library variants (CYRK_B200_LIB)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
import bench
for wl, batch in (("C5", 8 << 20), ("C2", 256 << 10), ("C2full", 64 << 10), ("C3", 1 << 20), ("C3full", 1 << 20)):
    w = bench.WORKLOADS[wl]
    y0, args = bench.make_inputs(w, batch, 100)
    y0d = torch.from_numpy(y0).cuda(); ad = torch.from_numpy(np.ascontiguousarray(args)).cuda()
    common = dict(method=w["method"], rtol=w["rtol"], atol=w["atol"], store="final"); common.update(w.get("kwargs", {}))
    if common.get("events") == "c3":
        common["events"] = [c.Event("y0"), c.Event("y1", direction=-1), c.Event("t_minus_arg4", terminal=True)]
    res = None; best = 1e9
    for rep in range(4):
        res = c.cysolve_ivp_batch(w["rhs"], (0.0, 10.0), y0d, args=ad, solution_reuse=res, **common)
        best = min(best, res.total_ms)
    print(os.environ.get("CYRK_B200_LIB", "default").split("/")[-1], wl, f"total_ms={best:.3f}", flush=True)
    res.close()
