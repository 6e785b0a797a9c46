#!/usr/bin/env python
"""bench.py -- trajectory-solves/s of the batched adaptive-RK path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C5|C1|C2|C2full|C3|C3full|C4|C4simple|C4ref]
                    [--batch B] [--impl reference] [--no-cpu-baseline] [--no-extras]

A "step" is one pass of the hot path over one batch of synthetic initial conditions (the shard of this
rank).  Default workload = BASELINE.json configs[4] / the north-star target: Lotka-Volterra (n_y=2) RK45,
rtol=1e-7, atol=1e-9, t_span=(0,10), 8 Mi trajectories PER GPU (weak scaling), store='final'.

Printed JSON line (rank 0):
  value     whole-job solves/s, inputs resident in HBM when the timed region starts (CYRK_FLAG_DEVICE_IO),
            timed on the device with CUDA events on the launching stream, max over ranks
  e2e       the same metric through the public API with HOST (pinned) buffers: H2D of y0 and D2H of the
            per-trajectory results are inside the timed region
  roofline  algorithmic FP64 FLOP/s of the solve kernel (200 FLOP per RK45 attempt for n_y=2,
            attempts from the kernel's own counters) against the FP64 peak measured in this run by a
            dependent-free DFMA probe (MEASURED_PEAKS.json has no FP64 figure); the HBM side is reported next
            to it to show the path is not bandwidth bound
  cpu_baseline   the reference's own CPU path (oracle/_ref when present, else the C port) on a bounded sample
  parity    (N=1) the GPU path against that same CPU reference on the SAME sample -- >= 1 Mi trajectories for the
            headline workload, 128 Ki for C2 / C3, 512 for C4 (128 at n_y = 10 000) -- in both FP builds: status equality, share of
            trajectories whose (accepted, rejected, rows, nfev) agree, largest error in units of the
            10*rtol*|y| + 10*atol tolerance.  The run FAILS (exit 2) when a step-sequence agreement is below 0.999.
  workloads (N=1) the other BASELINE configs timed the same way, 3 steps each: value, e2e, roofline fraction, bound
  config.size_sweep       (N=1) the headline workload at 16 / 32 / 64 Mi trajectories on one GPU (two steps each)
  config.strong_scaling   (N>1) a fixed 8 Mi-trajectory job split over the N GPUs (the weak-scaling headline keeps
            8 Mi per GPU); config.ensemble: the optional NCCL ensemble reduction, checked against numpy on rank 0

--impl reference times the CPU path alone (all host threads), same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# DRAM traffic per launch of the dominant kernel from `ncu --set full` (dram__bytes_read.sum + dram__bytes_write.sum),
# same workload/batch as the bench line.  (algorithmic for C5: y0 16 B + start-up record 40 B in, 44 B out per
# trajectory = 839 MB)
NCU_TRAFFIC = {"C5": (890.1e6, "profiles/r02_final_ncu_c5_lv_rk45.txt")}

# F_attempt(m) = n_y*c_m + E_m*F_rhs + s   
C_M = {"RK23": 25, "RK45": 62, "DOP853": 167}
E_M = {"RK23": 3, "RK45": 6, "DOP853": 12}
S_CTRL = {"RK23": 16, "RK45": 22, "DOP853": 34}
# per interpolant built / per interpolated point, per component; DOP853 builds cost 3 more RHS evaluations
F_DENSE = {"RK23": 11, "RK45": 33, "DOP853": 155}
F_INTERP = {"RK23": 10, "RK45": 13, "DOP853": 15}
Q_ORDER = {"RK23": 3, "RK45": 4, "DOP853": 7}

WORKLOADS = {
    # name: rhs, method, n_y, F_rhs, rtol, atol, per-GPU batch, description
    "C5": dict(rhs="lotkavolterra", method="RK45", ny=2, f_rhs=9, rtol=1e-7, atol=1e-9, batch=8 * 1024 * 1024,
               descr="C5/C1: Lotka-Volterra (n_y=2) RK45 rtol=1e-7 atol=1e-9 t_span=(0,10), store=final"),
    "C1": dict(rhs="lotkavolterra", method="RK45", ny=2, f_rhs=9, rtol=1e-7, atol=1e-9, batch=1024 * 1024,
               descr="C1: Lotka-Volterra (n_y=2) RK45 rtol=1e-7 atol=1e-9 t_span=(0,10), 1 Mi ICs, store=final"),
    "C2": dict(rhs="lorenz", method="DOP853", ny=3, f_rhs=9, rtol=1e-9, atol=1e-11, batch=256 * 1024,
               descr="C2 (final state only): Lorenz (n_y=3) DOP853 rtol=1e-9 atol=1e-11, rho~U(20,35), store=final"),
    "C3": dict(rhs="pendulum", method="RK23", ny=2, f_rhs=10, rtol=1e-6, atol=1e-8, batch=1024 * 1024,
               descr="C3 (no events): pendulum (n_y=2) RK23 rtol=1e-6 atol=1e-8, store=final"),
    # the full feature mixes of configs[1] / configs[2]
    "C2full": dict(rhs="lorenz", method="DOP853", ny=3, f_rhs=9, rtol=1e-9, atol=1e-11, batch=64 * 1024,
                   kwargs=dict(t_eval=np.linspace(0.0, 10.0, 1000), dense_output=True, store="rows", expected_size=448),
                   descr="C2: Lorenz DOP853 rtol=1e-9 atol=1e-11, t_eval 1000 pts + dense_output kept on device"),
    "C3full": dict(rhs="pendulum_drag", method="RK23", ny=2, f_rhs=12, rtol=1e-6, atol=1e-8, batch=1024 * 1024,
                   kwargs=dict(num_extra=1, store="final", events="c3"),
                   descr="C3: damped pendulum RK23 rtol=1e-6 atol=1e-8, 1 extra output, events theta=0 / omega=0 (dir -1) / "
                         "terminal t = t_stop(traj) ~ U(2,12), store=final + event lists"),
    # configs[3]: CTA-per-trajectory path.  F_rhs per COMPONENT: largey 6 + 1 sin (counted as 1), largey_simple ~0.
    "C4": dict(rhs="largey", method="RK45", ny=4096, f_rhs=7 * 4096, rtol=1e-6, atol=1e-8, batch=4096,
               kwargs=dict(store="rows", expected_size=40),
               descr="C4: largey (n_y=4096) RK45 rtol=1e-6 atol=1e-8, 4096 trajectories, every accepted row stored (4 GB)"),
    "C4simple": dict(rhs="largey_simple", method="RK45", ny=4096, f_rhs=0, rtol=1e-6, atol=1e-8, batch=4096,
                     kwargs=dict(store="rows", expected_size=40),
                     descr="C4 memory-side demonstrator: largey_simple (n_y=4096, RHS ~ free) RK45, 4096 trajectories, rows stored"),
    # the reference's own largey size (cysolver_test.pyx:224-237, Performance/largey.py:8-29: n_y = 10 000, y0 = 100, r = -0.5)
    "C4ref": dict(rhs="largey", method="RK45", ny=10000, f_rhs=7 * 10000, rtol=1e-6, atol=1e-8, batch=2048,
                  kwargs=dict(store="final"),
                  descr="largey at the reference's own size (n_y=10000) RK45 rtol=1e-6 atol=1e-8, 2048 trajectories, store=final"),
}
# what the default N=1 run adds to the headline line: (workload, steps)
EXTRA_WORKLOADS = [("C1", 3), ("C2full", 3), ("C3full", 3), ("C4", 3), ("C4simple", 3), ("C4ref", 2)]
# BASELINE sizes of configs[1] / configs[2] (4 Mi trajectories).  C2's t_eval samples are 32 KB per trajectory on the host:
# the batch is the largest power of two whose rows fit 35 % of the host memory available right now (4 Mi needs 134 GB).
FULL_BATCH = {"C2full": 4 * 1024 * 1024, "C3full": 4 * 1024 * 1024}


def host_mem_available():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return int(ln.split()[1]) * 1024
    except Exception:
        pass
    return 64 << 30


def full_batch(name):
    """Per-GPU batch of a BASELINE config at its stated size, bounded by what this host can hold."""
    b = FULL_BATCH.get(name)
    if b is None:
        return WORKLOADS[name]["batch"]
    if name == "C2full":
        per_traj = 1000 * 4 * 8
        while b > 65536 and b * per_traj > 0.35 * host_mem_available():
            b //= 2
    return b
# parity samples next to the headline one: workload -> trajectories
PARITY_SAMPLES = {"C2full": 128 * 1024, "C3full": 128 * 1024, "C4": 512, "C4ref": 128}


def make_inputs(w, n, seed):
    """Synthetic inputs of SURVEY.md 8d (numpy default_rng / PCG64)."""
    rng = np.random.default_rng(seed)
    if w["rhs"] in ("largey", "largey_simple"):
        y0 = 100.0 * (1 + 0.01 * rng.uniform(-1, 1, (n, w["ny"])))
        args = np.array([-0.5, float(w["ny"])])
        return np.ascontiguousarray(y0), args
    if w["rhs"] == "lotkavolterra":
        y0 = np.stack([rng.uniform(5, 15, n), rng.uniform(2.5, 7.5, n)], 1)
        args = np.array([1.5, 1.0, 3.0, 1.0])
    elif w["rhs"] == "lorenz":
        y0 = np.tile([1.0, 0.0, 0.0], (n, 1))
        args = np.stack([np.full(n, 10.0), rng.uniform(20, 35, n), np.full(n, 8.0 / 3.0)], 1)
    elif w["rhs"] == "pendulum_drag":
        y0 = np.stack([rng.uniform(-1.5, 1.5, n), rng.uniform(-1, 1, n)], 1)
        args = np.stack([np.ones(n), np.ones(n), np.full(n, 9.81), np.full(n, 1.5), rng.uniform(2.0, 12.0, n)], 1)
    else:
        y0 = np.stack([rng.uniform(-1.5, 1.5, n), rng.uniform(-1, 1, n)], 1)
        args = np.array([1.0, 1.0, 9.81, 0.0, 0.0])
    return np.ascontiguousarray(y0), np.ascontiguousarray(args)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.lines = gpu, None, []
        self.lo, self.hi = 0, None

    def start(self):
        """Started BEFORE the warm-up steps: nvidia-smi needs up to a second before its first line on an 8-GPU box, and
        the timed region of the headline workload is ~0.25 s (both arms)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark_begin(self):
        self.lo = len(self.lines)

    def mark_end(self):
        self.hi = len(self.lines)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        window = self.lines[self.lo:self.hi if self.hi is not None else len(self.lines)]
        note = "samples taken between the first timed step of the device-resident arm and the last of the end-to-end arm"
        if not window:               # (a timed region shorter than one sampling period: the samples right around it)
            window = self.lines[max(self.lo - 2, 0):(self.hi or 0) + 3]
            note = "timed region shorter than one sampling period: nearest samples"
        sm, smax, reasons = [], [], set()
        for ln in window:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "note": note}


# ----------------------------------------------------------------------------------------------------------------
# CPU reference leg (test infrastructure: the only place bench.py executes anything under oracle/)
# ----------------------------------------------------------------------------------------------------------------
C3_EVENTS = [("y0", 0, None), ("y1", -1, None), ("t_minus_arg4", 0, 1)]


def cpu_reference_run(w, n, threads=None, seed=12345, keep_outputs=False, impl=None, fix_quirk11=False):
    """The reference's CPU path on `n` trajectories of the workload (prange pattern, all host threads)."""
    from oracle import oracle as o
    kind = "reference" if o.have_ref() else "port"
    if impl is None:
        impl = "ref" if kind == "reference" else "port"
    y0, args = make_inputs(w, n, seed)
    threads = threads or len(os.sched_getaffinity(0))
    kw = dict(w.get("kwargs", {}))
    extra = {}
    if kw.get("events") == "c3":
        extra["events"] = C3_EVENTS
    if "t_eval" in kw:
        extra["t_eval"] = kw["t_eval"]
    if kw.get("dense_output"):
        extra["dense_output"] = True
    out = o.solve_batch(w["rhs"], (0.0, 10.0), y0, w["method"], w["rtol"], w["atol"], args=args, impl=impl,
                        n_threads=threads, want_teval=keep_outputs, fix_quirk11=fix_quirk11, **extra)
    ok = bool(np.all((out["status"] == 1) | (out["status"] == 2)))
    r = dict(seconds=out["seconds"], n=n, threads=threads, kind=kind, ok=ok, steps=int(out["steps"].sum()))
    if keep_outputs:
        r["out"] = out
        r["inputs"] = (y0, args)
    return r


def run_reference_arm(a, w, rank, world):
    if rank != 0:
        return
    n = 65536 if w["ny"] <= 32 else 512
    for _ in range(a.warmup):
        cpu_reference_run(w, 4096 if w["ny"] <= 32 else 64)
    secs, kind, threads = 0.0, "port", 1
    for _ in range(a.steps):
        r = cpu_reference_run(w, n)
        secs += r["seconds"]; kind = r["kind"]; threads = r["threads"]
    value = a.steps * n / secs
    sample = f"{n} trajectories per step ({a.steps} steps) of the workload's generator, all on rank 0"
    line = {
        "impl": "reference", "metric": "trajectory-solves/s", "value": value, "unit": "solves/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": secs / a.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["descr"], "trajectories_per_step": n, "host_threads": threads},
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


_REAL_STDOUT = None


def _guard_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner with a plain
    printf whenever NCCL_DEBUG is VERSION or WARN, which NCCL_DEBUG_FILE does not redirect), so file descriptor 1 is
    pointed at stderr for the whole run and the JSON line goes to a private duplicate of the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def _log(msg):
    sys.stderr.write(f"[bench] {msg}\n"); sys.stderr.flush()


# ----------------------------------------------------------------------------------------------------------------
# GPU legs
# ----------------------------------------------------------------------------------------------------------------
def solve_kwargs(c, w, device):
    kw = dict(method=w["method"], rtol=w["rtol"], atol=w["atol"], store="final", device=device)
    kw.update(w.get("kwargs", {}))
    if kw.get("events") == "c3":
        kw["events"] = [c.Event("y0"), c.Event("y1", direction=-1), c.Event("t_minus_arg4", terminal=True)]
    return kw


def gpu_measure(c, torch, w, B, seed, local_rank, steps, warmup, barrier, sample_clocks=False):
    """Both arms of one workload on this rank: device-resident (CUDA events of the shim) and end to end (host pinned
    buffers through cysolve_ivp_batch, wall clock between barriers)."""
    dev = torch.device("cuda", local_rank)
    y0_h, args_h = make_inputs(w, B, seed)
    y0_pin = torch.from_numpy(y0_h).pin_memory()
    args_pin = torch.from_numpy(np.ascontiguousarray(args_h)).pin_memory()
    y0_d = y0_pin.to(dev)
    args_d = args_pin.to(dev)
    common = solve_kwargs(c, w, local_rank)

    # ---- device-resident arm (value)
    sampler = ClockSampler(local_rank) if sample_clocks else None
    if sampler:
        sampler.start()
    res = None
    for _ in range(max(warmup, 3)):
        res = c.cysolve_ivp_batch(w["rhs"], (0.0, 10.0), y0_d, args=args_d, solution_reuse=res, **common)
    barrier()
    if sampler:
        sampler.mark_begin()
    dev_ms, ker_ms, launches = 0.0, 0.0, 0
    t0 = time.perf_counter()
    for _ in range(steps):
        res = c.cysolve_ivp_batch(w["rhs"], (0.0, 10.0), y0_d, args=args_d, solution_reuse=res, **common)
        dev_ms += res.total_ms           # CUDA events on the shim's launching stream, whole call
        ker_ms += res.kernel_ms          # CUDA events around the solve kernel(s) only
        launches += res.launches
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    res.close()

    # ---- end-to-end arm (host pinned buffers through the public API; H2D + D2H inside the timed region)
    res = None
    y0_np, args_np = y0_pin.numpy(), args_pin.numpy()
    for _ in range(2):
        res = c.cysolve_ivp_batch(w["rhs"], (0.0, 10.0), y0_np, args=args_np, solution_reuse=res, **common)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        res = c.cysolve_ivp_batch(w["rhs"], (0.0, 10.0), y0_np, args=args_np, solution_reuse=res, **common)
        _ = float(res.y_last[0, 0])      # touch the result on the host
    barrier()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = None
    if sampler:
        sampler.mark_end()
        clocks = sampler.stop()
    m = dict(B=B, dev_ms=dev_ms, ker_ms=ker_ms, launches=launches, wall_ms=wall_ms, e2e_wall_ms=e2e_wall_ms, clocks=clocks,
             h2d=int(res.h2d_bytes), d2h=int(res.d2h_bytes), e2e_launches=int(res.launches))
    # counters (fetched from the device arena on demand -- after the timed region)
    m["ok"] = bool(((res.status == 1) | (res.status == 2)).all())
    m["steps_acc"] = int(res.steps_taken.sum(dtype=np.int64))
    m["attempts"] = m["steps_acc"] + int(res.n_rejected.sum(dtype=np.int64))
    m["rows_saved"] = int(res.size.sum(dtype=np.int64))
    m["result"] = res                    # caller closes
    return m


def roofline_of(w, m, steps, fp64_peak, hbm_peak):
    """Algorithmic FLOP and bytes of one launch over the measured kernel time of this rank."""
    B = m["B"]
    f_attempt = w["ny"] * C_M[w["method"]] + E_M[w["method"]] * w["f_rhs"] + S_CTRL[w["method"]]
    flops = m["attempts"] * f_attempt
    kw_w = w.get("kwargs", {})
    if kw_w.get("dense_output"):       # one interpolant per accepted step
        flops += m["steps_acc"] * (w["ny"] * F_DENSE[w["method"]] + (3 * w["f_rhs"] if w["method"] == "DOP853" else 0))
    if "t_eval" in kw_w:               # one interpolated point per saved row
        flops += m["rows_saved"] * w["ny"] * F_INTERP[w["method"]]
    ker_ms = m["ker_ms"] / steps
    achieved_tf = flops / (ker_ms * 1e-3) / 1e12
    io_bytes = B * (w["ny"] * 8 + 5 * 4 + 8 + w["ny"] * 8)      # y0 in; counters, t_last, y_last out
    if kw_w.get("store") == "rows" and "t_eval" not in kw_w:
        io_bytes += (m["steps_acc"] + B) * (w["ny"] + 1) * 8      # every accepted row (t, y) of this rank
    if kw_w.get("store") == "rows" and "t_eval" in kw_w:
        io_bytes += m["rows_saved"] * (w["ny"] + 1) * 8           # every t_eval sample (t, y)
    if kw_w.get("dense_output"):                                  # stored interpolants: t_now + [y_old | Q] per step
        io_bytes += (m["steps_acc"] + B) * (1 + w["ny"] * (Q_ORDER[w["method"]] + 1)) * 8
    hbm_gbs = io_bytes / (ker_ms * 1e-3) / 1e9
    fp_frac = achieved_tf / fp64_peak if fp64_peak and fp64_peak > 0 else None
    hbm_frac = hbm_gbs / hbm_peak
    return dict(f_attempt=f_attempt, achieved_tf=achieved_tf, fp_frac=fp_frac, io_bytes=io_bytes, hbm_gbs=hbm_gbs,
                hbm_frac=hbm_frac, ker_ms=ker_ms)


def parity_check(c, w, name, n, local_rank, cpu=None):
    """GPU vs the CPU reference on the same `n` trajectories, both FP builds (north_star: step sequences agree for
    >= 99.9 % of trajectories, states within 10*rtol*|y| + 10*atol).  prange_test.pyx:87-121 is the reference's own
    batch-equality check."""
    if cpu is None:
        cpu = cpu_reference_run(w, n, keep_outputs=True)
    ref = cpu["out"]
    y0, args = cpu["inputs"]
    kw = solve_kwargs(c, w, local_rank)
    events = bool(w.get("kwargs", {}).get("events"))
    fixed = None
    if events:
        # on event termination with extra outputs the reference's final row is overrun garbage;
        # its correct value is what the oracle port computes with the overrun repaired (everything else is bit-equal)
        fixed = cpu_reference_run(w, n, keep_outputs=True, impl="port", fix_quirk11=True)["out"]
    out = {"n": int(n), "reference": cpu["kind"], "builds": {}}
    tol_of = lambda yr: 10.0 * w["rtol"] * np.abs(yr) + 10.0 * w["atol"]
    for build, strict in (("fma", False), ("strict", True)):
        r = c.cysolve_ivp_batch(w["rhs"], (0.0, 10.0), y0, args=args, strict_fp=strict, **kw)
        status_equal = bool(np.array_equal(r.status, ref["status"]))
        # accepted steps and saved rows always; without events nfev = 2 + evals_per_attempt * (accepted + rejected)
        # pins the rejected count too (with events + extra outputs nfev also counts Brent iterations: reported apart)
        nfev_same = np.asarray(r.nfev) == ref["nfev"]
        same = (r.steps_taken == ref["steps"]) & (r.size == ref["size"])
        if not events:
            same = same & nfev_same
        yr = ref["y_last"]
        if fixed is not None:
            yr = np.where((ref["status"] == 2)[:, None], fixed["y_last"], yr)
        err = np.abs(np.asarray(r.y_last) - yr) / tol_of(yr)
        finite = np.isfinite(yr).all(axis=1)
        d = {"status_equal": status_equal, "step_seq_agree": float(same.mean()), "nfev_agree": float(nfev_same.mean()),
             "max_err_over_tol": float(np.nanmax(err[finite])) if finite.any() else None}
        if "teval_y" in ref and r.y_rows is not None:
            te = ref["teval_y"]
            got = np.asarray(r.y_rows)[:, :te.shape[1], :]
            okm = np.isfinite(te)
            e2 = np.abs(got - te) / tol_of(te)
            d["teval_max_err_over_tol"] = float(np.max(e2[okm]))
            half = te.shape[1] // 2          # t <= 5: before the chaotic amplification (Lorenz) sets in
            d["teval_first_half_max_err_over_tol"] = float(np.max(e2[:, :half][okm[:, :half]]))
        out["builds"][build] = d
        r.close()
    out["status_equal"] = all(b["status_equal"] for b in out["builds"].values())
    out["step_seq_agree"] = min(b["step_seq_agree"] for b in out["builds"].values())
    out["max_err_over_tol"] = max(b["max_err_over_tol"] for b in out["builds"].values())
    return out


def main():
    _guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C5", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override trajectories per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the parity / workloads / strong-scaling / ensemble extras")
    a = ap.parse_args()
    w = dict(WORKLOADS[a.workload])
    if a.batch:
        w["batch"] = a.batch
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if a.impl == "reference":
        run_reference_arm(a, w, rank, world)
        return

    import torch
    import cyrk_b200 as c
    if not torch.cuda.is_available() or c.lib().cyrk_b200_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    numa_cores = None
    if world > 1:
        # one process per GPU: keep this rank (and, by first touch, its pinned host buffers) on the GPU's NUMA node
        from cyrk_b200.sharding import bind_to_device_numa
        numa_cores = bind_to_device_numa(local_rank)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    B = w["batch"]
    steps = a.steps
    fp64_peak = c.lib().cyrk_b200_measure_fp64_tflops(local_rank)
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm_peak = 6650.0
    # contiguous shard of the global IC set: rank r owns [r*B, (r+1)*B) -> generated with seed 100 + r
    m = gpu_measure(c, torch, w, B, 100 + rank, local_rank, steps, a.warmup, barrier, sample_clocks=True)
    res = m.pop("result")

    # ---- optional ensemble reduction over NCCL (the only collective this path has)
    ensemble = None
    if dist is not None and not a.no_extras:
        from cyrk_b200.sharding import ensemble_reduce_from_sums
        s1, s2, cnt = res.ensemble_sums()                     # local half: fixed-order device reduction
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lat = []
        for _ in range(5):
            e0.record()
            mean, var, count = ensemble_reduce_from_sums(s1, s2, cnt, device=dev)      # one all-reduce of 2*n_dy+1 doubles
            e1.record(); e1.synchronize()
            lat.append(e0.elapsed_time(e1))
        t0 = time.perf_counter()
        _ = res.ensemble_sums()
        local_ms = (time.perf_counter() - t0) * 1e3
        # check against numpy on the gathered final states (rank 0)
        y_all = [torch.empty((B, res.num_dy), dtype=torch.float64, device=dev) for _ in range(world)] if rank == 0 else None
        dist.gather(torch.from_numpy(np.array(res.y_last)).to(dev), y_all, dst=0)
        if rank == 0:
            yy = torch.cat(y_all).cpu().numpy()
            ref_mean, ref_var = yy.mean(0), yy.var(0)
            ensemble = {"mean": mean.tolist(), "var": var.tolist(), "count": int(count),
                        "max_rel_err_vs_numpy": float(max(np.max(np.abs(mean - ref_mean) / np.abs(ref_mean)),
                                                          np.max(np.abs(var - ref_var) / np.abs(ref_var)))),
                        "allreduce_ms_median": float(np.median(lat)), "local_device_sums_ms": local_ms,
                        "note": "local sums: fixed-order two-stage device reduction (cyrk_b200_ensemble_sums); "
                                "then ONE NCCL all-reduce of 2*n_dy+1 doubles"}
    res.close()

    # ---- strong scaling: a fixed 8 Mi-trajectory job split over the ranks
    strong = None
    if dist is not None and not a.no_extras and a.workload in ("C5", "C1"):
        total = 8 * 1024 * 1024
        ms = gpu_measure(c, torch, w, total // world, 1000 + rank, local_rank, steps, a.warmup, barrier)
        ms.pop("result").close()
        s_dev, s_e2e = max_over_ranks([ms["dev_ms"], ms["e2e_wall_ms"]])
        strong = {"total_trajectories": total, "trajectories_per_gpu": total // world,
                  "value": steps * total / (s_dev * 1e-3), "ms_per_step": s_dev / steps,
                  "e2e_value": steps * total / (s_e2e * 1e-3), "e2e_ms_per_step": s_e2e / steps}

    # ---- size sweep of the headline workload on one GPU (C5 = the C1 generator at 8, 16, 32, 64 Mi trajectories)
    sweep = None
    if dist is None and not a.no_extras and a.workload == "C5":
        sweep = {}
        for mi in (16, 32, 64):
            _log(f"size sweep: {mi} Mi trajectories")
            nb = mi * 1024 * 1024
            ms = gpu_measure(c, torch, w, nb, 100 + int(np.log2(mi // 8)), local_rank, 2, 3, barrier)
            ms.pop("result").close()
            sweep[f"{mi}Mi"] = {"value": 2 * nb / (ms["dev_ms"] * 1e-3), "ms_per_step": ms["dev_ms"] / 2,
                                "e2e": 2 * nb / (ms["e2e_wall_ms"] * 1e-3), "e2e_ms_per_step": ms["e2e_wall_ms"] / 2,
                                "all_success": bool(ms["ok"])}
            torch.cuda.empty_cache()

    # ---- max over ranks
    dev_ms, ker_ms, e2e_wall_ms, wall_ms = max_over_ranks([m["dev_ms"], m["ker_ms"], m["e2e_wall_ms"], m["wall_ms"]])
    cnt = torch.tensor([m["attempts"], m["steps_acc"], m["launches"] + m["e2e_launches"] * steps, int(m["ok"])], dtype=torch.int64, device=dev)
    if dist is not None:
        mn = cnt[3:4].clone()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        cnt[3] = mn[0]
    attempts_all, steps_all, launches_all, ok_all = (int(v) for v in cnt.tolist())

    exit_code = 0
    if rank == 0:
        total = B * world
        value = steps * total / (dev_ms * 1e-3)
        e2e_value = steps * total / (e2e_wall_ms * 1e-3)
        rf = roofline_of(w, m, steps, fp64_peak, hbm_peak)       # per-GPU kernel roofline (rank 0's counters; equal shards)
        traffic = NCU_TRAFFIC.get(a.workload) if B == WORKLOADS[a.workload]["batch"] else None
        line = {
            "metric": "trajectory-solves/s", "value": value, "unit": "solves/s", "n_gpus": world,
            "steps": steps, "warmup": max(a.warmup, 3), "ms_per_step": dev_ms / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["descr"], "trajectories_per_gpu": B, "t_span": [0.0, 10.0],
                       "l2": "inputs+outputs per step (%.0f MB) exceed the 126 MB L2" % (rf["io_bytes"] / 1e6),
                       "rk_attempts_per_step": attempts_all, "rk_steps_per_s": steps_all / (dev_ms / steps * 1e-3),
                       "all_success": bool(ok_all), "wall_ms_per_step": wall_ms / steps,
                       "host_binding": ("rank pinned to its GPU's NUMA-local cores (%d on rank 0)" % len(numa_cores)) if numa_cores
                                       else "none"},
            "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                    "ms_per_step": e2e_wall_ms / steps,
                    "note": "per trajectory: y0 in; y_last + status + steps_taken out; size / n_rejected / nfev / t_last stay "
                            "in the device arena until their getter is called (pull-based getters like the reference's)"},
            "gpu_launches": launches_all,
            "clocks": m["clocks"],
            "roofline": {"bound": "fp64", "achieved": rf["achieved_tf"], "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": rf["fp_frac"],
                         "traffic": traffic[0] if traffic else None,
                         "traffic_unit": ("bytes per launch (ncu, %s)" % traffic[1]) if traffic else None,
                         "peak_source": "in-run dependent-free DFMA probe (cyrk_b200_measure_fp64_tflops); "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "flop_per_attempt": rf["f_attempt"], "kernel_ms_per_launch": rf["ker_ms"],
                         "hbm": {"achieved_gbs": rf["hbm_gbs"], "peak_gbs": hbm_peak, "frac": rf["hbm_frac"]}},
        }
        if strong:
            line["config"]["strong_scaling"] = strong
        if sweep:
            line["config"]["size_sweep"] = sweep
        if ensemble:
            line["config"]["ensemble"] = ensemble
        if world == 1 and not a.no_cpu_baseline:
            # bounded sample: calibrate on 4096 trajectories, then aim for ~10 s of CPU work (>= 1 Mi for the headline
            # workload so that the parity check below covers at least the batch of BASELINE config 1)
            n_cal = 4096 if w["ny"] <= 32 else 64
            cal = cpu_reference_run(w, n_cal)
            rate = n_cal / max(cal["seconds"], 1e-6)
            n = int(min(max(rate * 10.0, 2 * n_cal), (1 << 21) if w["ny"] <= 32 else 8192))   # (large n_y: bound the host copy of y0)
            if a.workload in ("C5", "C1") and not a.no_extras:
                n = max(n, 1 << 20)
            r = cpu_reference_run(w, n, keep_outputs=not a.no_extras)
            line["cpu_baseline"] = {"value": r["n"] / r["seconds"], "unit": "solves/s", "cores": r["threads"],
                                    "kind": r["kind"],
                                    "sample": f"{r['n']} trajectories of the same generator (seed 12345), "
                                              f"{r['seconds']:.1f} s on {r['threads']} host threads"}
            if not a.no_extras:
                _log("parity: GPU vs the CPU reference on the cpu_baseline sample")
                par = {a.workload: parity_check(c, w, a.workload, n, local_rank, cpu=r)}
                del r
                if a.workload == "C5":
                    for name, pn in PARITY_SAMPLES.items():
                        _log(f"parity: {name} on {pn} trajectories")
                        par[name] = parity_check(c, dict(WORKLOADS[name]), name, pn, local_rank)
                worst = min(p["step_seq_agree"] for p in par.values())
                line["parity"] = {"n": par[a.workload]["n"], "status_equal": all(p["status_equal"] for p in par.values()),
                                  "step_seq_agree": worst,
                                  "max_err_over_tol": par[a.workload]["max_err_over_tol"],
                                  "bar": "status equal; accepted steps, saved rows and (via nfev) rejected attempts equal for >= 99.9 % of trajectories; "
                                         "|y - y_ref| <= 10*rtol*|y_ref| + 10*atol (max_err_over_tol <= 1)",
                                  "per_workload": par}
                if worst < 0.999 or not line["parity"]["status_equal"]:
                    exit_code = 2
        if world == 1 and not a.no_extras and a.workload == "C5":
            wl = {}
            for name, k in EXTRA_WORKLOADS:
                ww = dict(WORKLOADS[name])
                ww["batch"] = full_batch(name)
                _log(f"workload {name}, {ww['batch']} trajectories")
                try:
                    mm = gpu_measure(c, torch, ww, ww["batch"], 100, local_rank, k, 3, barrier)
                    mm.pop("result").close()
                    rr = roofline_of(ww, mm, k, fp64_peak, hbm_peak)
                    bound = "fp64" if (rr["fp_frac"] or 0) >= rr["hbm_frac"] else "hbm"
                    wl[name] = {"batch": ww["batch"], "steps": k, "value": k * ww["batch"] / (mm["dev_ms"] * 1e-3),
                                "e2e": k * ww["batch"] / (mm["e2e_wall_ms"] * 1e-3), "unit": "solves/s",
                                "ms_per_step": mm["dev_ms"] / k, "e2e_ms_per_step": mm["e2e_wall_ms"] / k,
                                "kernel_ms_per_step": rr["ker_ms"], "bound": bound,
                                "roofline": {"frac": rr["fp_frac"] if bound == "fp64" else rr["hbm_frac"],
                                             "fp64_frac": rr["fp_frac"], "hbm_frac": rr["hbm_frac"],
                                             "achieved_tflops": rr["achieved_tf"], "achieved_gbs": rr["hbm_gbs"]},
                                "h2d_bytes_per_step": mm["h2d"], "d2h_bytes_per_step": mm["d2h"], "all_success": mm["ok"],
                                "descr": ww["descr"]}
                except Exception as e:       # a workload that cannot run here must not take the headline line with it
                    wl[name] = {"error": str(e)[:300]}
            line["workloads"] = wl
        _emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if exit_code:
        sys.exit(exit_code)


if __name__ == "__main__":
    main()
