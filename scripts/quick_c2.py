"""C2full / C3full kernel timing with a block-size sweep (register-heavy ROWS / FULL variants)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
import bench
for wl in ("C2full", "C3full"):
    w = bench.WORKLOADS[wl]
    B = w["batch"]
    y0, args = bench.make_inputs(w, B, 100)
    y0d = torch.from_numpy(y0).cuda(); ad = torch.from_numpy(np.ascontiguousarray(args)).cuda()
    common = dict(method=w["method"], rtol=w["rtol"], atol=w["atol"], store="final"); common.update(w.get("kwargs", {}))
    if common.get("events") == "c3":
        common["events"] = [c.Event("y0"), c.Event("y1", direction=-1), c.Event("t_minus_arg4", terminal=True)]
    for block in (0, 128, 96):
        c.lib().cyrk_b200_set_option(b"block", block)
        res = None; best = 1e9
        for rep in range(4):
            res = c.cysolve_ivp_batch(w["rhs"], (0.0, 10.0), y0d, args=ad, solution_reuse=res, **common)
            best = min(best, res.total_ms)
        print(wl, "block", block, f"total_ms={best:.3f}", flush=True)
        res.close()
