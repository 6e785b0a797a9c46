import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cyrk_b200 as c
rng = np.random.default_rng(2)
B = 1 << 20
y0 = torch.from_numpy(np.stack([rng.uniform(-1.5, 1.5, B), rng.uniform(-1, 1, B)], 1)).cuda()
args = torch.from_numpy(np.stack([np.ones(B), np.ones(B), np.full(B, 9.81), np.full(B, 1.5), rng.uniform(2.0, 12.0, B)], 1)).cuda()
ts = torch.tensor([0.0, 10.0], dtype=torch.float64, device="cuda")
def run(tag, **kw):
    res = None; best = 1e9
    for rep in range(3):
        res = c.cysolve_ivp_batch("pendulum_drag", ts, y0, method="RK23", args=args, rtol=1e-6, atol=1e-8, num_extra=1, store="final", solution_reuse=res, **kw)
        best = min(best, res.kernel_ms)
    st = torch.as_tensor(res.steps_taken, device="cuda").sum().item(); nr = torch.as_tensor(res.n_rejected, device="cuda").sum().item()
    import hashlib
    dg = hashlib.md5(torch.as_tensor(res.y_last, device="cuda").cpu().numpy().tobytes() + torch.as_tensor(res.nfev, device="cuda").cpu().numpy().tobytes()).hexdigest()[:12]
    print(f"{tag:40s} kernel_ms={best:8.2f} attempts/traj={(st+nr)/B:7.1f} solves/s={B/best*1e3:.3e} digest={dg}", flush=True)
    res.close()
rb = int(os.environ.get("RB", 8)); c.lib().cyrk_b200_set_option(b"root_batch", rb); print("root_batch", rb)
run("no events (LEAN)")
run("1 event that never triggers (FULL)", events=[c.Event("y0_gt_50")])
run("1 non-terminal event theta", events=[c.Event("y0")])
run("3 events, terminal t_stop (C3)", events=[c.Event("y0"), c.Event("y1", direction=-1), c.Event("t_minus_arg4", terminal=True)])
run("t_eval 21 pts (ROWS)", t_eval=np.linspace(0, 10, 21))
run("dense_output (FULL)", dense_output=True, expected_size=2048) if False else None
