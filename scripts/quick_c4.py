"""C4 one-liner: largey n_y=QN (4096), QB (4096) trajectories, RK45 / RK23, store=final; kernel time and a checksum."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
rng = np.random.default_rng(3)
n = int(os.environ.get("QN", 4096))
B = int(os.environ.get("QB", 4096))
y0d = torch.from_numpy(100.0 * (1 + 0.01 * rng.uniform(-1, 1, (B, n)))).cuda()
ts = torch.tensor([0.0, 10.0], dtype=torch.float64, device="cuda")
ar = torch.tensor([-0.5, float(n)], dtype=torch.float64, device="cuda")
c.lib().cyrk_b200_set_option(b"cta2_wide", int(os.environ.get("WIDE", 0)))
for rhs, method in (("largey", "RK45"), ("largey", "RK23"), ("largey_simple", "RK45")):
    res = None
    best = 1e9
    for rep in range(4):
        res = c.cysolve_ivp_batch(rhs, ts, y0d, method=method, args=ar, rtol=1e-6, atol=1e-8, store="final", solution_reuse=res)
        best = min(best, res.kernel_ms)
    st = torch.as_tensor(res.steps_taken, device="cuda").cpu().numpy().astype(np.int64)
    yl = torch.as_tensor(res.y_last, device="cuda").cpu().numpy()
    ok = bool((torch.as_tensor(res.status, device="cuda") == 1).all())
    print(f"{rhs:14s} {method} n={n} B={B} kernel_ms={best:.3f} steps={int(st.sum())} ok={ok} sum_y={yl.sum():.15e}", flush=True)
    res.close()
