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


class _DevicePointer:
    """A raw device allocation as a zero-copy torch tensor (CUDA array interface, bytes)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def wrap_pointer(ptr, nbytes, device=None):
    """uint8 tensor over `nbytes` at `ptr` without a copy: device memory when `device` is a CUDA device, else host
    memory (the gloo tests)."""
    import torch
    if device is not None and torch.device(device).type == "cuda":
        return torch.as_tensor(_DevicePointer(ptr, nbytes), device=device)
    import ctypes
    buf = (ctypes.c_uint8 * int(nbytes)).from_address(int(ptr))
    return torch.from_numpy(np.frombuffer(buf, dtype=np.uint8))


def broadcast_field_in_place(ptr, nbytes, src=0, device=None):
    """The one collective of a scene update, device to device: rank `src`'s resident field (its own buffer, e.g.
    smplgpu_distance_field_dev_ptr) is broadcast straight into every other rank's reserved buffer
    (smplgpu_reserve_distance_field) -- no host hop, no staging tensor.  Returns the wrapped tensor."""
    import torch.distributed as dist
    t = wrap_pointer(ptr, nbytes, device)
    dist.broadcast(t, src=src)
    return t


def gather_counts(local_value, device=None):
    """Sum of a per-rank scalar (units processed) and max of a per-rank time, as bench.py reports them."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(local_value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
