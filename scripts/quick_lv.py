"""C5 kernel timing sweep: refill_min x block (prep kernel on), 8 Mi Lotka-Volterra trajectories by default."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cyrk_b200 as c
rng = np.random.default_rng(0)
B = int(os.environ.get('QB', 1 << 20))
import torch
y0 = torch.from_numpy(np.stack([rng.uniform(5, 15, B), rng.uniform(2.5, 7.5, B)], 1)).cuda()
args_d = torch.tensor([1.5, 1, 3, 1], dtype=torch.float64, device='cuda'); ts_d = torch.tensor([0., 10.], dtype=torch.float64, device='cuda')
cfgs = [(1, 1, 128)] if not os.environ.get("SWEEP") else [(1, r, b) for r in (1, 2, 3, 4, 6) for b in (128, 256)]
if os.environ.get("SWEEP") == "prep":
    cfgs = [(pr, r, 128) for pr in (0, 1) for r in (1, 2, 4)]
for prep, refill, block in cfgs:
    c.lib().cyrk_b200_set_option(b'block', block)
    c.lib().cyrk_b200_set_option(b"refill_min", refill)
    try: c.lib().cyrk_b200_set_option(b"prep", prep)
    except Exception: pass
    best = 1e9
    for rep in range(4):
        r = c.cysolve_ivp_batch('lotkavolterra', ts_d, y0, method='RK45', args=args_d, rtol=1e-7, atol=1e-9, store='final')
        best = min(best, r.kernel_ms)
    att = int(torch.as_tensor(r.steps_taken, device='cuda').sum().item()) + int(torch.as_tensor(r.n_rejected, device='cuda').sum().item())
    print(os.environ.get("CYRK_B200_LIB", "default"), f"prep={prep} refill={refill} block={block} kernel_ms={best:.3f} attempts={att}", flush=True)
