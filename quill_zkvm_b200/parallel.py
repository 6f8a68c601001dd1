"""Multi-GPU plumbing: one process per GPU, torch.distributed only for bootstrap and timing barriers.

The data path has no torch collective.  The path's two exchanges run inside libquill_b200.so on the
context's stream over NCCL: an all-gather of one XYZZ partial sum per rank (MSM) and an all-gather of deg+1 partial
sums per round (sumcheck).  This module shards index ranges and hands the ncclUniqueId from rank 0 to the others.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Contiguous range [lo, hi) of rank `rank` when n items are split over `world` ranks (n % world == 0 for tables)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def table_shard_range(num_vars: int, rank: int, world: int):
    """Split a 2^num_vars table by its top log2(world) variables: rank g holds [g*2^m, (g+1)*2^m), m = num_vars - log2 world.
    Pairs (2p, 2p+1) stay local because the sumcheck binds variable 0 first (hyperplonk/src/piops/sumcheck.rs:51-57)."""
    assert world & (world - 1) == 0, "rank count must be a power of two"
    n = 1 << num_vars
    assert n >= world, "fewer table entries than ranks"
    per = n // world
    return rank * per, (rank + 1) * per


def broadcast_bytes(buf: np.ndarray, src: int = 0) -> np.ndarray:
    """Broadcast a small uint8 array from `src` over the default process group (gloo or nccl)."""
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(buf, dtype=np.uint8).copy())
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=src)
    return t.cpu().numpy()


def init_comm(ctx) -> None:
    """Create the library's NCCL communicator across the default process group."""
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    uid = ctx.comm_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)
    uid = broadcast_bytes(uid, 0)
    ctx.comm_init(uid, rank, world)
