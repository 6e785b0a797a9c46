#!/bin/bash
P=${1:-r02_s56}
O=gpurun_out
mkdir -p $O
python bench.py --workload C3full --no-extras --no-cpu-baseline --steps 3 > $O/${P}_bench_c3full.json 2> $O/${P}_bench_c3full.err; tail -c 700 $O/${P}_bench_c3full.json; echo
python scripts/quick_c3full_e2e.py 2>&1 | tee $O/${P}_c3full_e2e.log
python - <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import cyrk_b200 as c, bench
w = bench.WORKLOADS["C3full"]; B = 1 << 22
y0, args = bench.make_inputs(w, B, 100)
y0p, ap = torch.from_numpy(y0).pin_memory().numpy(), torch.from_numpy(np.ascontiguousarray(args)).pin_memory().numpy()
kw = bench.solve_kwargs(c, w, 0)
for taper in (1, 0, 1, 0):
    c.lib().cyrk_b200_set_option(b"stream_taper", taper)
    res = None
    for _ in range(2):
        res = c.cysolve_ivp_batch(w["rhs"], (0.0, 10.0), y0p, args=ap, solution_reuse=res, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        res = c.cysolve_ivp_batch(w["rhs"], (0.0, 10.0), y0p, args=ap, solution_reuse=res, **kw)
        _ = float(res.y_last[0, 0])
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"taper={taper}: per-call wall ms {['%.1f' % t for t in ts]} kernels {res.kernel_ms:.1f} total_ms {res.total_ms:.1f}", flush=True)
    res.close()
PY
