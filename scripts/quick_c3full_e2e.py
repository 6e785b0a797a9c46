"""C3full at 4 Mi trajectories end to end (host pinned buffers): wall per call vs the chunk size of the streamed solve."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
import bench
w = bench.WORKLOADS["C3full"]
B = int(os.environ.get("QB", 1 << 22))
y0, args = bench.make_inputs(w, B, 100)
y0p, ap = torch.from_numpy(y0).pin_memory().numpy(), torch.from_numpy(np.ascontiguousarray(args)).pin_memory().numpy()
kw = bench.solve_kwargs(c, w, 0)
for chunk, taper in ((0, 1), (0, 0), (1 << 20, 0), (1 << 21, 0)):
    c.lib().cyrk_b200_set_option(b"stream_chunk", chunk); c.lib().cyrk_b200_set_option(b"stream_taper", taper)
    res = None
    for _ in range(2):
        res = c.cysolve_ivp_batch(w["rhs"], (0.0, 10.0), y0p, args=ap, solution_reuse=res, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        res = c.cysolve_ivp_batch(w["rhs"], (0.0, 10.0), y0p, args=ap, solution_reuse=res, **kw)
        _ = float(res.y_last[0, 0])
    wall = (time.perf_counter() - t0) * 1e3 / n
    print(f"stream_chunk={chunk:8d} taper={taper}: wall {wall:.1f} ms per call ({B / wall * 1e3:.3e} solves/s), kernels {res.kernel_ms:.1f} ms, launches {res.launches}, d2h {res.d2h_bytes / 1e9:.2f} GB", flush=True)
    res.close()
c.lib().cyrk_b200_set_option(b"stream_chunk", 0); c.lib().cyrk_b200_set_option(b"stream_taper", 1)
