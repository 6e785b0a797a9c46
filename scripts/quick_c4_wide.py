"""C4 (largey n_y=4096, 4096 trajectories): 512 threads x 8 components vs 256 threads x 16 components (option cta2_wide)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
rng = np.random.default_rng(3)
n = int(os.environ.get("QN", 4096))
B = int(os.environ.get("QB", 4096))
y0 = 100.0 * (1 + 0.01 * rng.uniform(-1, 1, (B, n)))
y0d = torch.from_numpy(y0).cuda()
ts = torch.tensor([0.0, 10.0], dtype=torch.float64, device="cuda")
ar = torch.tensor([-0.5, float(n)], dtype=torch.float64, device="cuda")
ref = {}
for wide in (0, 1):
    c.lib().cyrk_b200_set_option(b"cta2_wide", wide)
    for rhs in ("largey", "largey_simple"):
        for method in ("RK45", "RK23"):
            for store in ("final", "rows"):
                res = None
                best = 1e9
                for rep in range(3):
                    res = c.cysolve_ivp_batch(rhs, ts, y0d, method=method, args=ar, rtol=1e-6, atol=1e-8, store=store, expected_size=40 if method == "RK45" else 200, solution_reuse=res)
                    best = min(best, res.kernel_ms)
                st = torch.as_tensor(res.steps_taken, device="cuda").cpu().numpy().astype(np.int64)
                nr = torch.as_tensor(res.n_rejected, device="cuda").cpu().numpy().astype(np.int64)
                ok = bool((torch.as_tensor(res.status, device="cuda") == 1).all())
                yl = torch.as_tensor(res.y_last, device="cuda").cpu().numpy().copy()
                key = (rhs, method, store)
                d = ""
                if key in ref:
                    d = f" max_rel_diff_vs_narrow={np.max(np.abs(yl - ref[key][0]) / np.abs(ref[key][0]).clip(1e-300)):.2e} steps_equal={bool((st == ref[key][1]).all())}"
                else:
                    ref[key] = (yl, st)
                att = int(st.sum() + nr.sum())
                print(f"wide={wide} {rhs:14s} {method} store={store:5s} kernel_ms={best:.2f} attempts={att} ok={ok} launches={res.launches}{d}", flush=True)
                res.close()
