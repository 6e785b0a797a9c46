"""CPU: host-side multi-GPU logic with the gloo backend, world_size 2 (the N>1 path of bench.py /
cyrk_b200.sharding).  The solve itself is stubbed with the CPU oracle -- only the plumbing is under test."""
import os
import socket
import sys

import numpy as np
import pytest

from cyrk_b200.sharding import shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_the_batch():
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from types import SimpleNamespace
    from cyrk_b200.sharding import ensemble_reduce, solve_sharded
    from oracle import oracle as o
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def solve_fn(diffeq, t_span, y0, args=None, **kw):
        out = o.solve_batch(diffeq, t_span, y0, kw["method"], kw["rtol"], kw["atol"], args=args, impl="port", n_threads=1)
        return SimpleNamespace(status=out["status"], steps_taken=out["steps"], t_last=out["t_last"], y_last=out["y_last"])

    rng = np.random.default_rng(7)
    n = 37
    y0 = np.stack([rng.uniform(5, 15, n), rng.uniform(2.5, 7.5, n)], 1)
    args = np.stack([rng.uniform(1.4, 1.6, n), np.ones(n), np.full(n, 3.0), np.ones(n)], 1)   # per-trajectory args
    local, gathered = solve_sharded(solve_fn, "lotkavolterra", (0.0, 5.0), y0, args=args, method="RK45", rtol=1e-6, atol=1e-8)
    mean, var, cnt = ensemble_reduce(local.y_last, local.status == 1)
    if rank == 0:
        q.put(dict(gathered=gathered, mean=mean, var=var, cnt=cnt, y0=y0, args=args))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_sharded_solve_and_ensemble_reduce(oracle):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = oracle.solve_batch("lotkavolterra", (0.0, 5.0), res["y0"], "RK45", 1e-6, 1e-8, args=res["args"], impl="port", n_threads=1)
    g = res["gathered"]
    assert np.array_equal(g["status"], full["status"])
    assert np.array_equal(g["steps_taken"], full["steps"])
    assert np.array_equal(g["y_last"], full["y_last"])            # sharding must not change a single bit
    assert res["cnt"] == 37
    assert np.allclose(res["mean"], full["y_last"].mean(0), rtol=1e-12)
    assert np.allclose(res["var"], full["y_last"].var(0), rtol=1e-9)
