"""Multi-GPU execution, one process per GPU: orbits are independent, so the orbit index is cut into contiguous
slices, one per rank, with NO collective on the data path.  (The other form -- ONE process,
one call, all devices -- lives inside the C ABI: ``gb_launch.n_devices``, ``gala_b200.set_devices``; it uses the same
``shard_bounds`` rule for orbits and deals mock-stream particles in groups of 128 rows, ``capi.cu:Deal``.)
``torch.distributed`` is only plumbing: rank/world discovery and, when the caller wants the full
result on every rank or on rank 0, a gather of the per-rank (6, [ntimes,] n_r) blocks.
"""
from __future__ import annotations

import numpy as np

__all__ = ["shard_bounds", "shard", "gather_orbits", "integrate_sharded"]


def shard_bounds(N: int, world: int):
    """Contiguous slices [lo, hi) of range(N) for each rank; sizes differ by at most one."""
    base, rem = divmod(N, world)
    lo = 0
    out = []
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def shard(w0, rank: int, world: int):
    """This rank's columns of a (6, N) array (a view; made contiguous by the integrators)."""
    lo, hi = shard_bounds(w0.shape[-1], world)[rank]
    return w0[..., lo:hi]


def gather_orbits(local, N: int, dst=None, group=None):
    """Gather per-rank blocks (..., n_r) along the last axis into (..., N).  ``local`` is a numpy array
    (gloo/CPU object gather) or a torch tensor (nccl: padded all_gather, results stay on the device).
    With ``dst`` set only that rank returns the array (others return None)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = shard_bounds(N, world)
    is_np = isinstance(local, np.ndarray)
    t = torch.as_tensor(local) if is_np else local
    nmax = max(hi - lo for lo, hi in bounds)
    pad = torch.zeros(t.shape[:-1] + (nmax,), dtype=t.dtype, device=t.device)
    pad[..., : t.shape[-1]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    if dst is not None and rank != dst:
        return None
    full = torch.cat([b[..., : hi - lo] for b, (lo, hi) in zip(bufs, bounds)], dim=-1)
    return full.numpy() if is_np else full


def integrate_sharded(fn, hamiltonian, w0, t, gather=True, dst=None, **kw):
    """Run one of the boundary functions (``leapfrog_integrate_hamiltonian`` ...) on this rank's slice
    of ``w0`` (6, N) and optionally gather the result.  Returns ``(t_out, w)``."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(), dist.get_rank()
    else:
        world, rank = 1, 0
    N = w0.shape[-1]
    mine = shard(w0, rank, world)
    tt, w = fn(hamiltonian, mine, t, **kw)[:2]
    if gather and world > 1:
        w = gather_orbits(w, N, dst=dst)
    return tt, w
