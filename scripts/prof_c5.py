"""Two C5 solves (Lotka-Volterra RK45, 8 Mi trajectories, device-resident, final state) for ncu (skip the first)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
import bench
w = bench.WORKLOADS["C5"]
y0, args = bench.make_inputs(w, w["batch"], 100)
y0d = torch.from_numpy(y0).cuda(); ad = torch.from_numpy(args).cuda()
res = None
for _ in range(2):
    res = c.cysolve_ivp_batch("lotkavolterra", (0.0, 10.0), y0d, args=ad, method="RK45", rtol=1e-7, atol=1e-9, store="final", solution_reuse=res)
print("kernel_ms", res.kernel_ms)
