"""Multi-GPU plumbing: trajectories are independent (Documentation/Parallelization.md:9-12), so the batch is
sharded by contiguous ranges of initial conditions, one process per GPU, with NO collective on the solve
path.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only to gather per-trajectory results
or for the optional ensemble reduction (sum y, sum y^2, count -> one all-reduce of 2*n_dy+1 doubles).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of trajectories owned by `rank`; the first n % world ranks get one extra."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def _slice_per_traj(x, lo, hi, n):
    if x is None:
        return None
    a = np.asarray(x)
    if a.ndim >= 2 and a.shape[0] == n:
        return a[lo:hi]
    return x


def solve_sharded(solve_fn: Callable, diffeq, t_span, y0, args=None, group=None, gather: bool = True, **kw):
    """Each rank solves y0[lo:hi] with `solve_fn` (default use: cyrk_b200.cysolve_ivp_batch bound to its local
    GPU).  Returns (local_result, gathered) where `gathered` (rank 0 only, when `gather`) is a dict of the
    concatenated per-trajectory arrays status / steps_taken / t_last / y_last in global order."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    y0 = np.asarray(y0, dtype=np.float64)
    n = y0.shape[0]
    lo, hi = shard_bounds(n, rank, world)
    local = solve_fn(diffeq, _slice_per_traj(t_span, lo, hi, n), y0[lo:hi], args=_slice_per_traj(args, lo, hi, n), **kw)
    gathered = None
    if gather:
        payload = dict(lo=lo, hi=hi,
                       status=np.asarray(local.status).copy(), steps_taken=np.asarray(local.steps_taken).copy(),
                       t_last=np.asarray(local.t_last).copy(), y_last=np.asarray(local.y_last).copy())
        if world == 1:
            parts = [payload]
        else:
            parts = [None] * world if rank == 0 else None
            dist.gather_object(payload, parts, dst=0, group=group)
        if rank == 0:
            parts.sort(key=lambda p: p["lo"])
            gathered = {k: np.concatenate([p[k] for p in parts]) for k in ("status", "steps_taken", "t_last", "y_last")}
    return local, gathered


def ensemble_reduce_from_sums(sum_y, sum_y2, count, group=None, device=None):
    """Second half of the ensemble reduction: all-reduce(SUM) of the (2*n_dy + 1) local sums -- e.g. the ones
    BatchResult.ensemble_sums() computed on the device -- and turn them into (mean, var, count)."""
    import torch
    import torch.distributed as dist
    buf = np.concatenate([np.asarray(sum_y, dtype=np.float64), np.asarray(sum_y2, dtype=np.float64), [float(count)]])
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    out = t.cpu().numpy()
    ndy = (out.size - 1) // 2
    cnt = out[-1]
    mean = out[:ndy] / cnt
    return mean, out[ndy:2 * ndy] / cnt - mean * mean, int(cnt)


def ensemble_reduce(y_last, ok, group=None, device=None):
    """Ensemble mean / variance of the final states over ALL ranks: local sums, then one all-reduce(SUM) of
    (2*n_dy + 1) float64 values.  Returns (mean[n_dy], var[n_dy], count)."""
    import torch
    import torch.distributed as dist
    y = np.asarray(y_last, dtype=np.float64)
    ok = np.asarray(ok, dtype=bool)
    ys = y[ok]
    buf = np.concatenate([ys.sum(0), (ys * ys).sum(0), [float(ok.sum())]])
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    out = t.cpu().numpy()
    ndy = y.shape[1]
    cnt = out[-1]
    mean = out[:ndy] / cnt
    var = out[ndy:2 * ndy] / cnt - mean * mean
    return mean, var, int(cnt)


def solve_on_devices(diffeq, t_span, y0, devices, args=None, **kw):
    """Drive several GPUs from ONE process: one host thread (and one stream, owned by its result handle) per
    device, contiguous shards, no communication.  The C entry points release nothing Python-side (ctypes drops the
    GIL for the duration of a call), so the shards run concurrently.  Returns a list of (lo, hi, BatchResult).

    This is the in-process alternative to one-process-per-GPU (bench.py / torchrun); both use the same sharding."""
    import threading
    from .solver import cysolve_ivp_batch
    y0 = np.asarray(y0, dtype=np.float64)
    n = y0.shape[0]
    world = len(devices)
    out = [None] * world
    errs = [None] * world

    def work(r):
        lo, hi = shard_bounds(n, r, world)
        try:
            res = cysolve_ivp_batch(diffeq, _slice_per_traj(t_span, lo, hi, n), y0[lo:hi],
                                    args=_slice_per_traj(args, lo, hi, n), device=int(devices[r]), **kw)
            out[r] = (lo, hi, res)
        except Exception as e:          # re-raised in the caller's thread
            errs[r] = e

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errs:
        if e is not None:
            raise e
    return out


def bind_to_device_numa(device: int) -> Optional[List[int]]:
    """Pin the calling process to the CPU cores NVML reports as local to `device` (its PCIe root / NUMA node).

    One process per GPU moves its shard through pinned host memory; with first-touch page placement the pinned buffers
    then live on the GPU's own NUMA node, so the H2D / D2H copies of eight ranks do not cross the socket interconnect.
    Call it BEFORE allocating pinned memory.  Returns the core list, or None when NVML / the topology is unavailable
    (the process is then left as it was)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cores = [64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None
