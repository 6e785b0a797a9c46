"""C1 (1 Mi LV trajectories, final state): resident CTAs per SM vs kernel time -- fewer resident warps shorten the drain
phase (each lane's last trajectory) but lower the steady-state throughput."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
rng = np.random.default_rng(0)
for B in (1 << 20, 1 << 23):
    y0 = torch.from_numpy(np.stack([rng.uniform(5, 15, B), rng.uniform(2.5, 7.5, B)], 1)).cuda()
    args_d = torch.tensor([1.5, 1, 3, 1], dtype=torch.float64, device='cuda'); ts_d = torch.tensor([0., 10.], dtype=torch.float64, device='cuda')
    for prep in (-1, 1, 0):
        for ctas in (0, 8, 7, 6, 5, 4):
            if prep != -1 and ctas not in (0, 6): continue
            c.lib().cyrk_b200_set_option(b"ctas_per_sm", ctas); c.lib().cyrk_b200_set_option(b"prep", prep)
            best = 1e9; res = None
            for rep in range(5):
                res = c.cysolve_ivp_batch('lotkavolterra', ts_d, y0, method='RK45', args=args_d, rtol=1e-7, atol=1e-9, store='final', solution_reuse=res)
                best = min(best, res.kernel_ms)
            print(f"B={B} prep={prep} ctas_per_sm={ctas} kernel_ms={best:.3f} total_ms={res.total_ms:.3f} launches={res.launches}", flush=True)
            res.close()
c.lib().cyrk_b200_set_option(b"ctas_per_sm", 0); c.lib().cyrk_b200_set_option(b"prep", -1)
