"""C4 timing: largey n_y=4096, RK45, 4096 trajectories (CTA-per-trajectory path) + largey_simple (memory-bound demo)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cyrk_b200 as c
rng = np.random.default_rng(3)
n = 4096
B = int(os.environ.get("QB", 4096))
y0 = 100.0 * (1 + 0.01 * rng.uniform(-1, 1, (B, n)))
import torch
y0d = torch.from_numpy(y0).cuda()
ts = torch.tensor([0.0, 10.0], dtype=torch.float64, device="cuda")
ar = torch.tensor([-0.5, float(n)], dtype=torch.float64, device="cuda")
import itertools
KERNEL = int(os.environ.get("CTA_KERNEL", 0)); c.lib().cyrk_b200_set_option(b"cta_kernel", KERNEL)
cfgs = [(-1, 0)] + [(v, b) for v in (9, 7, 4, 3, 2, 0) for b in (1024, 512)] if os.environ.get("SWEEP") else [(-1, 0)]
for (sv, blk), rhs in itertools.product(cfgs, ("largey", "largey_simple")):
    c.lib().cyrk_b200_set_option(b"cta_smem_vecs", sv); c.lib().cyrk_b200_set_option(b"cta_block", blk)
    print("smem_vecs", sv, "block", blk, end="  ")
    for store in (("rows",) if os.environ.get("SWEEP") else ("final", "rows")):
        res = None
        best = 1e9
        for rep in range(3):
            res = c.cysolve_ivp_batch(rhs, ts, y0d, method="RK45", args=ar, rtol=1e-6, atol=1e-8, store=store, expected_size=40, solution_reuse=res)
            best = min(best, res.kernel_ms)
        st = torch.as_tensor(res.steps_taken, device="cuda").cpu().numpy().astype(np.int64)
        nr = torch.as_tensor(res.n_rejected, device="cuda").cpu().numpy().astype(np.int64)
        sz = torch.as_tensor(res.size, device="cuda").cpu().numpy().astype(np.int64)
        ok = bool((torch.as_tensor(res.status, device="cuda") == 1).all())
        att = int(st.sum() + nr.sum())
        io = B * n * 8 + (int(sz.sum()) * (n + 1) * 8 if store == "rows" else B * n * 8)
        flop = att * (n * 62 + 6 * n * 7 + 22)
        print(f"{rhs:14s} store={store:5s} B={B} kernel_ms={best:.2f} solves/s={B/best*1e3:.3e} attempts={att} steps/traj={st.mean():.1f} "
              f"io_GB={io/1e9:.2f} hbm_GBs={io/best/1e6:.0f} alg_TFLOPs={flop/best/1e9:.2f} ok={ok} launches={res.launches}", flush=True)
        res.close()
