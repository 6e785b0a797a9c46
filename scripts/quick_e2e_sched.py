"""End-to-end (host pinned buffers) final-state timing: head/body/tail schedule (pipe_sched 1) against equal chunks with tapered
ends (pipe_sched 0), per batch size and pipe_chunk."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
import bench
w = bench.WORKLOADS["C5"]
L = c.lib()
kw = dict(method="RK45", rtol=1e-7, atol=1e-9, store="final")
for B in (1 << 18, 1 << 20, 1 << 21, 1 << 22, 1 << 23):
    y0, args = bench.make_inputs(w, B, 100)
    y0p = torch.from_numpy(y0).pin_memory().numpy(); ap = torch.from_numpy(np.ascontiguousarray(args)).pin_memory().numpy()
    y0d = torch.from_numpy(y0).cuda(); ad = torch.from_numpy(args).cuda()
    res = None
    for _ in range(3):
        res = c.cysolve_ivp_batch("lotkavolterra", (0.0, 10.0), y0d, args=ad, solution_reuse=res, **kw)
    dev_ms = min(c.cysolve_ivp_batch("lotkavolterra", (0.0, 10.0), y0d, args=ad, **kw).total_ms for _ in range(3))
    res.close()
    print(f"B={B} device-resident total_ms={dev_ms:.3f}", flush=True)
    for rep in range(2):
        for sched, chunk in ((0, 1 << 21), (1, 1 << 21), (1, 3 << 20), (1, 1 << 22)):
            if sched == 1 and chunk != (1 << 21) and B < (1 << 22): continue
            L.cyrk_b200_set_option(b"pipe_sched", sched); L.cyrk_b200_set_option(b"pipe_chunk", chunk)
            res = None
            for _ in range(2):
                res = c.cysolve_ivp_batch("lotkavolterra", (0.0, 10.0), y0p, args=ap, solution_reuse=res, **kw)
            torch.cuda.synchronize(); n = 8; best = 1e9; tot = 0.0
            for _ in range(n):
                t0 = time.perf_counter()
                res = c.cysolve_ivp_batch("lotkavolterra", (0.0, 10.0), y0p, args=ap, solution_reuse=res, **kw)
                _ = float(res.y_last[0, 0])
                dt = (time.perf_counter() - t0) * 1e3
                best = min(best, dt); tot += dt
            print(f"  B={B} pipe_sched={sched} pipe_chunk={chunk}: wall_ms mean={tot / n:.3f} best={best:.3f} launches={res.launches} "
                  f"-> {B / (tot / n) * 1e3:.3e} solves/s ({dev_ms / (tot / n):.3f} of device-resident)", flush=True)
            res.close()
    L.cyrk_b200_set_option(b"pipe_sched", 1); L.cyrk_b200_set_option(b"pipe_chunk", 1 << 21)
