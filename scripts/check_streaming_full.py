"""BASELINE-size streamed solves (C2full, C3full) through the public API with host buffers: timing + bitwise equality of a
64 Ki subset with the single-arena path (trajectories are independent, so a subset solved alone must give the same bits)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cyrk_b200 as c
import bench

def run(name, B, sub=65536):
    w = dict(bench.WORKLOADS[name])
    y0, args = bench.make_inputs(w, B, 100)
    kw = bench.solve_kwargs(c, w, 0)
    res = None
    for it in range(2):
        t0 = time.perf_counter()
        res = c.cysolve_ivp_batch(w["rhs"], (0.0, 10.0), y0, args=args, solution_reuse=res, **kw)
        dt = time.perf_counter() - t0
        print(f"{name} B={B} call {it}: wall {dt*1e3:.1f} ms, device total {res.total_ms:.1f} ms, kernels {res.kernel_ms:.1f} ms, launches {res.launches}, "
              f"d2h {res.d2h_bytes/1e9:.2f} GB -> {B/dt:.3e} solves/s", flush=True)
    c.lib().cyrk_b200_set_option(b"stream", 0)
    ref = c.cysolve_ivp_batch(w["rhs"], (0.0, 10.0), y0[:sub], args=args[:sub] if np.ndim(args) == 2 else args, **kw)
    c.lib().cyrk_b200_set_option(b"stream", 1)
    ok = True
    for nm in ("status", "size", "steps_taken", "n_rejected", "nfev", "t_last", "y_last"):
        same = np.array_equal(np.asarray(getattr(res, nm))[:sub], np.asarray(getattr(ref, nm)), equal_nan=True)
        ok &= same
        if not same: print("  MISMATCH", nm)
    if res.row_capacity:
        n = np.asarray(ref.size)
        a, b = np.asarray(res.y_rows)[:sub], np.asarray(ref.y_rows)
        m = np.arange(a.shape[1])[None, :] < n[:, None]
        same = np.array_equal(a[:, :b.shape[1]][m[:, :b.shape[1]]], b[m[:, :b.shape[1]]])
        ok &= same; print("  rows equal:", same)
    if res.num_events:
        E = res.num_events
        oa, ob = np.asarray(res.event_offsets), np.asarray(ref.event_offsets)
        same = np.array_equal(oa[:sub * E + 1], ob) and np.array_equal(np.asarray(res.event_t)[:int(ob[-1])], np.asarray(ref.event_t)) \
            and np.array_equal(np.asarray(res.event_y)[:int(ob[-1])], np.asarray(ref.event_y))
        ok &= same; print("  events equal:", same, "total hits", res.event_total)
    if name == "C2full":
        tq = np.linspace(0.5, 9.5, 7)
        qa, ca = res.call(tq); qb, cb = ref.call(tq)
        same = np.array_equal(qa[:sub], qb) and np.array_equal(ca[:sub], cb)
        ok &= same; print("  sol(t) equal:", same)
    print(f"{name} B={B}: subset of {sub} bitwise equal to the single-arena path: {ok}", flush=True)
    res.close(); ref.close()

if __name__ == "__main__":
    run("C3full", int(os.environ.get("C3B", 4 << 20)))
    run("C2full", int(os.environ.get("C2B", 1 << 20)))
