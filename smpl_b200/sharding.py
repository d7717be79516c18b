"""Multi-GPU plumbing: the path shards with no exchange step.

States, edges and planning queries are independent, so rank r of W simply takes a slice of them; the only
collective is ONE broadcast of the scene (the uint16 squared-distance field; every rank derives its BFS walls
from it) per scene update.  Works with any torch.distributed backend: NCCL over NVLink on the GPU box, gloo in
the CPU tests.
"""
import numpy as np


def contiguous_shard(n, rank, world):
    """Sweep API: rank r takes the r-th contiguous slice of ceil(n / world) items.  Returns (begin, end)."""
    per = (n + world - 1) // world
    b = min(n, rank * per)
    return b, min(n, b + per)


def round_robin_shard(n, rank, world):
    """Planning queries: query i goes to rank i mod world (hard and easy queries interleave)."""
    return np.arange(rank, n, world)


def broadcast_distance_field(d2, dims, src=0, device=None):
    """One broadcast of the distance field.  `d2`: uint16 array on the source rank (ignored elsewhere).
    Returns a torch.uint8 tensor holding the 2 * nx*ny*nz bytes of the uint16 field on `device` (bytes, because
    every backend broadcasts uint8)."""
    import torch
    import torch.distributed as dist
    n = int(dims[0]) * int(dims[1]) * int(dims[2])
    t = torch.empty(2 * n, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        t.copy_(torch.from_numpy(np.ascontiguousarray(d2, dtype=np.uint16).view(np.uint8).reshape(-1)))
    dist.broadcast(t, src=src)
    return t


def gather_counts(local_value, device=None):
    """Sum of a per-rank scalar (units processed) and max of a per-rank time, as bench.py reports them."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(local_value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
