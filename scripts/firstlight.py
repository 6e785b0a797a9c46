"""First-light check on a B200: parity vs the CPU oracle on small batches + a rough timing."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cyrk_b200 as c
from oracle import oracle as o

def compare(name, rhs, method, y0, args, rtol, atol, strict, **kw):
    r = c.cysolve_ivp_batch(rhs, (0, 10), y0, method=method, args=args, rtol=rtol, atol=atol, strict_fp=strict, store='rows', **kw)
    ob = o.solve_batch(rhs, (0, 10), y0, method, rtol, atol, args=args, impl='port', fix_quirk11=True, **{k: v for k, v in kw.items() if k in ('t_eval','dense_output')})
    same_status = np.array_equal(r.status, ob['status'])
    same_steps = (r.steps_taken == ob['steps']).mean()
    same_size = (r.size == ob['size']).mean()
    same_nfev = (r.nfev == ob['nfev']).mean()
    err = np.abs(r.y_last - ob['y_last'])
    tol = 10 * rtol * np.abs(ob['y_last']) + 10 * atol
    bitexact = (r.y_last == ob['y_last']).all(axis=1).mean()
    print(f"{name:28s} strict={strict} status_ok={same_status} steps_eq={same_steps:.4f} size_eq={same_size:.4f} nfev_eq={same_nfev:.4f} "
          f"within_tol={(err<=tol).all()} max_err={err.max():.3e} bitexact={bitexact:.4f} kernel_ms={r.kernel_ms:.3f}", flush=True)

rng = np.random.default_rng(0)
B = 4096
y0 = np.stack([rng.uniform(5, 15, B), rng.uniform(2.5, 7.5, B)], 1)
print("devices", c.lib().cyrk_b200_device_count(), flush=True)
for strict in (True, False):
    for m in ('RK23', 'RK45', 'DOP853'):
        compare(f'LV {m}', 'lotkavolterra', m, y0, [1.5, 1, 3, 1], 1e-7, 1e-9, strict)
    rho = rng.uniform(20, 35, 1024)
    a = np.stack([np.full(1024, 10.0), rho, np.full(1024, 8 / 3)], 1)
    compare('Lorenz DOP853', 'lorenz', 'DOP853', np.tile([1.0, 0, 0], (1024, 1)), a, 1e-9, 1e-11, strict)
    compare('Lorenz DOP853 teval+dense', 'lorenz', 'DOP853', np.tile([1.0, 0, 0], (256, 1)), a[:256], 1e-9, 1e-11, strict, t_eval=np.linspace(0, 10, 100), dense_output=True)
    yp = np.stack([rng.uniform(-1.5, 1.5, 1024), rng.uniform(-1, 1, 1024)], 1)
    compare('Pendulum RK23', 'pendulum', 'RK23', yp, [1, 1, 9.81], 1e-6, 1e-8, strict)

print("fp64 probe TFLOP/s:", c.lib().cyrk_b200_measure_fp64_tflops(0), flush=True)
# rough throughput
for B in (1 << 20,):
    y0 = np.stack([rng.uniform(5, 15, B), rng.uniform(2.5, 7.5, B)], 1)
    for prep, refill in ((1, 1), (1, 2), (1, 4), (1, 8), (0, 1), (0, 4), (0, 8), (0, 32)):
        c.lib().cyrk_b200_set_option(b"refill_min", refill)
        c.lib().cyrk_b200_set_option(b"prep", prep)
        for rep in range(2):
            r = c.cysolve_ivp_batch('lotkavolterra', (0, 10), y0, method='RK45', args=[1.5, 1, 3, 1], rtol=1e-7, atol=1e-9, store='final')
        att = int(r.steps_taken.sum()) + int(r.n_rejected.sum())
        print(f"B={B} prep={prep} refill_min={refill} kernel_ms={r.kernel_ms:.3f} total_ms={r.total_ms:.3f} solves/s={B/r.kernel_ms*1e3:.3e} attempts={att} "
              f"alg_TFLOPs={att*200/r.kernel_ms/1e9:.2f} ok={(r.status==1).all()}", flush=True)
