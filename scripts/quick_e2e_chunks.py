"""C5 end to end (8 Mi LV trajectories, host pinned buffers): wall per call vs pipe_chunk."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
import bench
w = bench.WORKLOADS["C5"]; L = c.lib()
kw = dict(method="RK45", rtol=1e-7, atol=1e-9, store="final")
for B in (1 << 23, 1 << 20):
    y0, args = bench.make_inputs(w, B, 100)
    y0p = torch.from_numpy(y0).pin_memory().numpy(); ap = torch.from_numpy(np.ascontiguousarray(args)).pin_memory().numpy()
    for chunk, taper in ((1 << 20, 1), (1 << 21, 1), (3 << 20, 1), (1 << 22, 1), (1 << 21, 0), (1 << 19, 1)):
        if chunk >= B and B < (1 << 23): continue
        L.cyrk_b200_set_option(b"pipe_chunk", chunk); L.cyrk_b200_set_option(b"pipe_taper", taper)
        res = None
        for _ in range(3):
            res = c.cysolve_ivp_batch("lotkavolterra", (0.0, 10.0), y0p, args=ap, solution_reuse=res, **kw)
        torch.cuda.synchronize(); n = 8; tot = 0.0; best = 1e9
        for _ in range(n):
            t0 = time.perf_counter()
            res = c.cysolve_ivp_batch("lotkavolterra", (0.0, 10.0), y0p, args=ap, solution_reuse=res, **kw)
            _ = float(res.y_last[0, 0])
            dt = (time.perf_counter() - t0) * 1e3; tot += dt; best = min(best, dt)
        print(f"B={B} pipe_chunk={chunk} taper={taper}: wall mean {tot / n:.3f} ms best {best:.3f} launches={res.launches} -> {B / (tot / n) * 1e3:.3e} solves/s", flush=True)
        res.close()
L.cyrk_b200_set_option(b"pipe_chunk", 1 << 21); L.cyrk_b200_set_option(b"pipe_taper", 1)
