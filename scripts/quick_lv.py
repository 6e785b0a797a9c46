import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cyrk_b200 as c
rng = np.random.default_rng(0)
B = int(os.environ.get('QB', 1 << 20))
y0 = np.stack([rng.uniform(5, 15, B), rng.uniform(2.5, 7.5, B)], 1)
for prep, refill in ((0, 1), (0, 2), (0, 3), (0, 4), (0, 6), (0, 8), (1, 1), (1, 2), (1, 4)):
    c.lib().cyrk_b200_set_option(b"refill_min", refill)
    try: c.lib().cyrk_b200_set_option(b"prep", prep)
    except Exception: pass
    best = 1e9
    for rep in range(4):
        r = c.cysolve_ivp_batch('lotkavolterra', (0, 10), y0, method='RK45', args=[1.5, 1, 3, 1], rtol=1e-7, atol=1e-9, store='final')
        best = min(best, r.kernel_ms)
    att = int(r.steps_taken.sum()) + int(r.n_rejected.sum())
    print(os.environ.get("CYRK_B200_LIB", "default"), f"prep={prep} refill={refill} kernel_ms={best:.3f} attempts={att}", flush=True)
