"""One C2full solve (Lorenz DOP853, t_eval 1000 + dense, 64 Ki) on the step-parallel path, for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
import bench
w = bench.WORKLOADS["C2full"]
B = w["batch"]
y0, args = bench.make_inputs(w, B, 100)
y0d = torch.from_numpy(y0).cuda(); ad = torch.from_numpy(np.ascontiguousarray(args)).cuda()
kw = dict(method=w["method"], rtol=w["rtol"], atol=w["atol"]); kw.update(w.get("kwargs", {}))
res = None
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    res = c.cysolve_ivp_batch(w["rhs"], (0.0, 10.0), y0d, args=ad, solution_reuse=res, **kw)
print("kernel_ms", res.kernel_ms, "launches", res.launches)
