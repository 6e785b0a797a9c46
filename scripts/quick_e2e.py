"""End-to-end (host pinned buffers) C5 timing vs the pipeline chunk size."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
import bench
w = bench.WORKLOADS["C5"]; B = w["batch"]
y0, args = bench.make_inputs(w, B, 100)
y0p = torch.from_numpy(y0).pin_memory().numpy(); ap = torch.from_numpy(np.ascontiguousarray(args)).pin_memory().numpy()
for chunk, taper in ((1 << 20, 0), (1 << 20, 1), (1 << 19, 1), (3 << 19, 1)):
    c.lib().cyrk_b200_set_option(b"pipe_chunk", chunk); c.lib().cyrk_b200_set_option(b"pipe_taper", taper)
    res = None
    for _ in range(2):
        res = c.cysolve_ivp_batch("lotkavolterra", (0.0, 10.0), y0p, args=ap, method="RK45", rtol=1e-7, atol=1e-9, store="final", solution_reuse=res)
    torch.cuda.synchronize(); t0 = time.perf_counter(); n = 5; dev = 0.0
    for _ in range(n):
        res = c.cysolve_ivp_batch("lotkavolterra", (0.0, 10.0), y0p, args=ap, method="RK45", rtol=1e-7, atol=1e-9, store="final", solution_reuse=res)
        dev += res.total_ms; _ = float(res.y_last[0, 0])
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / n * 1e3
    print(f"chunk={chunk} taper={taper} wall_ms={wall:.2f} device_total_ms={dev / n:.2f} launches={res.launches}", flush=True)
    res.close()
