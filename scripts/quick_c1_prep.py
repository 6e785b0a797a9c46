"""Final-state LV solves: start-up kernel (prep) + immediate refill vs inline set-up, by batch size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
rng = np.random.default_rng(0)
args_d = torch.tensor([1.5, 1, 3, 1], dtype=torch.float64, device='cuda'); ts_d = torch.tensor([0., 10.], dtype=torch.float64, device='cuda')
for B in (1 << 14, 1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 21):
    y0 = torch.from_numpy(np.stack([rng.uniform(5, 15, B), rng.uniform(2.5, 7.5, B)], 1)).cuda()
    out = []
    for prep, refill in ((0, 4), (0, 1), (1, 1), (1, 2)):
        c.lib().cyrk_b200_set_option(b"prep", prep); c.lib().cyrk_b200_set_option(b"refill_min", refill)
        best = 1e9; bt = 1e9; res = None
        for rep in range(6):
            res = c.cysolve_ivp_batch('lotkavolterra', ts_d, y0, method='RK45', args=args_d, rtol=1e-7, atol=1e-9, store='final', solution_reuse=res)
            best = min(best, res.kernel_ms); bt = min(bt, res.total_ms)
        out.append(f"prep={prep}/refill={refill}: {best:.3f} ({bt:.3f})")
        res.close()
    print(f"B={B:8d}  " + "   ".join(out), flush=True)
c.lib().cyrk_b200_set_option(b"prep", -1); c.lib().cyrk_b200_set_option(b"refill_min", 0)
