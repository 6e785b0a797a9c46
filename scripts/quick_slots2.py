"""Two-slot A/B: kernel times and output digests of the final-state solves of C5 / C1 (LV RK45 / RK23) and the pendulum (RK23),
per library (CYRK_B200_LIB).  Digests of two libraries must be equal: the slots run the one-slot kernel's arithmetic."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cyrk_b200 as c

tag = os.path.basename(os.environ.get("CYRK_B200_LIB", "default"))
rng = np.random.default_rng(0)


def host(x):
    return np.ascontiguousarray(torch.as_tensor(x, device='cuda').cpu().numpy())


def digest(r):
    h = hashlib.md5()
    for name in ("y_last", "steps_taken", "status", "n_rejected", "nfev"):
        h.update(host(getattr(r, name)).tobytes())
    return h.hexdigest()[:12]


def run(label, problem, method, B, y0, args, rtol, atol, t_end, opts=()):
    y0_d = torch.from_numpy(y0).cuda()
    args_d = torch.from_numpy(np.asarray(args, dtype=np.float64)).cuda()
    ts_d = torch.tensor([0., t_end], dtype=torch.float64, device='cuda')
    for k, v in opts: c.lib().cyrk_b200_set_option(k, v)
    best = 1e9
    for rep in range(4):
        r = c.cysolve_ivp_batch(problem, ts_d, y0_d, method=method, args=args_d, rtol=rtol, atol=atol, store='final')
        best = min(best, r.kernel_ms)
    att = int(host(r.steps_taken).astype(np.int64).sum() + host(r.n_rejected).astype(np.int64).sum())
    print(f"{tag} {label} B={B} opts={opts} kernel_ms={best:.3f} attempts={att} digest={digest(r)}", flush=True)


def lv(B): return np.stack([rng.uniform(5, 15, B), rng.uniform(2.5, 7.5, B)], 1)
def pend(B): return np.stack([rng.uniform(-1.5, 1.5, B), rng.uniform(-1, 1, B)], 1)


if os.environ.get("PREP"): c.lib().cyrk_b200_set_option(b'prep', int(os.environ["PREP"]))
for Bs in (1, 77, 4099):      # slots that never hold a trajectory, partially filled warps
    run("LV/RK45 small", 'lotkavolterra', 'RK45', Bs, lv(Bs), [1.5, 1, 3, 1], 1e-7, 1e-9, 10.0)
sizes = [int(x) for x in os.environ.get("QB", "1048576,8388608").split(",")]
for B in sizes:
    y = lv(B)
    run("LV/RK45", 'lotkavolterra', 'RK45', B, y, [1.5, 1, 3, 1], 1e-7, 1e-9, 10.0)
    if os.environ.get("SWEEP"):
        for blk, ctas in ((256, 0), (64, 0), (128, 3), (128, 2)):
            run("LV/RK45", 'lotkavolterra', 'RK45', B, y, [1.5, 1, 3, 1], 1e-7, 1e-9, 10.0, ((b'block', blk), (b'ctas_per_sm', ctas)))
        c.lib().cyrk_b200_set_option(b'block', 0); c.lib().cyrk_b200_set_option(b'ctas_per_sm', 0)
        for rm in (2, 4, 8):
            run("LV/RK45", 'lotkavolterra', 'RK45', B, y, [1.5, 1, 3, 1], 1e-7, 1e-9, 10.0, ((b'refill_min', rm),))
        c.lib().cyrk_b200_set_option(b'refill_min', 0)
B = sizes[0]
run("LV/RK23", 'lotkavolterra', 'RK23', B, lv(B), [1.5, 1, 3, 1], 1e-6, 1e-8, 10.0)
run("LV/DOP853", 'lotkavolterra', 'DOP853', B, lv(B), [1.5, 1, 3, 1], 1e-9, 1e-11, 10.0)
if os.environ.get("NOPEND"): sys.exit(0)
run("pendulum/RK23", 'pendulum', 'RK23', B, pend(B), [1.0, 1.0, 9.81, 0.0, 0.0], 1e-6, 1e-8, 10.0)
run("pendulum/RK45", 'pendulum', 'RK45', B, pend(B), [1.0, 1.0, 9.81, 0.0, 0.0], 1e-6, 1e-8, 10.0)
