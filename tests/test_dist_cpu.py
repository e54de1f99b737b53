"""`-m "not gpu"`: the N>1 path (orbit-index sharding + gather) under gloo, world_size 2."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from gala_b200.dist import gather_orbits, integrate_sharded, shard, shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N = 11
        w0 = np.arange(6 * N, dtype=np.float64).reshape(6, N)
        t = np.arange(4.0)

        # stand-in for a boundary function: "integrates" by adding t[-1] (no GPU in this container)
        def fake_integrate(H, w, tt, save_all=1):
            out = np.repeat(w[:, None, :], len(tt), axis=1) + tt[None, :, None]
            return tt, (out if save_all else out[:, -1])
        mine = shard(w0, rank, world)
        lo, hi = shard_bounds(N, world)[rank]
        assert mine.shape == (6, hi - lo)
        tt, full = integrate_sharded(fake_integrate, None, w0, t, save_all=1)
        expect = np.repeat(w0[:, None, :], 4, axis=1) + t[None, :, None]
        ok = full.shape == (6, 4, N) and np.array_equal(full, expect)
        tt, last = integrate_sharded(fake_integrate, None, w0, t, save_all=0, dst=0)
        ok = ok and ((rank == 0 and np.array_equal(last, expect[:, -1])) or (rank != 0 and last is None))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_sharded_integration_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]
