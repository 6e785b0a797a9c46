"""C5 kernel with the shared args in the parameter block (VAR_LEAN_SHARED) vs per-lane registers (option shared_args = 0)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
rng = np.random.default_rng(0)
ts = torch.tensor([0., 10.], dtype=torch.float64, device="cuda")
ad = torch.tensor([1.5, 1, 3, 1], dtype=torch.float64, device="cuda")
for B in (1 << 23, 1 << 20):
    y0 = torch.from_numpy(np.stack([rng.uniform(5, 15, B), rng.uniform(2.5, 7.5, B)], 1)).cuda()
    ref = None
    for rep in range(2):
        for sh in (0, 1):
            c.lib().cyrk_b200_set_option(b"shared_args", sh)
            best = 1e9
            for k in range(3):
                r = c.cysolve_ivp_batch("lotkavolterra", ts, y0, method="RK45", args=ad, rtol=1e-7, atol=1e-9, store="final")
                best = min(best, r.kernel_ms)
            yl = torch.as_tensor(r.y_last, device="cuda").clone(); st = torch.as_tensor(r.steps_taken, device="cuda").clone()
            if ref is None: ref = (yl, st)
            same = bool(torch.equal(ref[0], yl) and torch.equal(ref[1], st))
            print(f"B={B} shared_args={sh} kernel_ms={best:.3f} bitwise_equal_to_first={same}", flush=True)
