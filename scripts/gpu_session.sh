#!/bin/bash
P=${1:-r02_s40}
O=gpurun_out
mkdir -p $O
python - <<PY 2>&1 | tee -a $O/${P}_c2full_refill.log
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import cyrk_b200 as c, bench
w = bench.WORKLOADS["C2full"]; B = 1 << 21
y0, args = bench.make_inputs(w, B, 100)
y0d, ad = torch.from_numpy(y0).cuda(), torch.from_numpy(args).cuda()
kw = bench.solve_kwargs(c, w, 0)
for prep, refill in ((-1, 0), (0, 1), (0, 2), (0, 4), (0, 8), (1, 4), (1, 1)):
    c.lib().cyrk_b200_set_option(b"prep", prep); c.lib().cyrk_b200_set_option(b"refill_min", refill)
    res = None; best = 1e9
    for _ in range(3):
        res = c.cysolve_ivp_batch(w["rhs"], (0.0, 10.0), y0d, args=ad, solution_reuse=res, **kw)
        best = min(best, res.kernel_ms)
    print(f"prep={prep} refill_min={refill} C2full B={B} kernel_ms={best:.3f} launches={res.launches}", flush=True)
    res.close()
PY
timeout 900 python -m pytest tests/test_gpu_streaming.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
