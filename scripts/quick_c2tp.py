"""C2full (Lorenz DOP853, t_eval 1000 + dense, 64 Ki): per-trajectory kernel vs the step-parallel path, device-resident."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
import bench
w = bench.WORKLOADS["C2full"]
for B in (w["batch"], 4 * w["batch"]):
    y0, args = bench.make_inputs(w, B, 100)
    y0d = torch.from_numpy(y0).cuda(); ad = torch.from_numpy(np.ascontiguousarray(args)).cuda()
    common = dict(method=w["method"], rtol=w["rtol"], atol=w["atol"]); common.update(w.get("kwargs", {}))
    for mode in ("t_eval+dense", "t_eval", "dense"):
        kw = dict(common)
        if mode == "t_eval": kw["dense_output"] = False
        if mode == "dense": kw["t_eval"] = None
        for tp in (0, 1):
            c.lib().cyrk_b200_set_option(b"two_pass", tp)
            res = None; best = 1e9; bk = 1e9
            for rep in range(4):
                res = c.cysolve_ivp_batch(w["rhs"], (0.0, 10.0), y0d, args=ad, solution_reuse=res, **kw)
                best = min(best, res.total_ms); bk = min(bk, res.kernel_ms)
            print(f"B={B} {mode:13s} two_pass={tp} total_ms={best:.3f} kernel_ms={bk:.3f} launches={res.launches}", flush=True)
            res.close()
c.lib().cyrk_b200_set_option(b"two_pass", 1)
