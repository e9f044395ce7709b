"""Multi-GPU plumbing: one process per GPU, contiguous slices of the batch per rank,
tree / obstacle buffers replicated with ONE broadcast when they change, no data-path collective.
torch.distributed is used only for that broadcast, the barrier and the max-over-ranks timing reduction;
the backend is NCCL on the GPU box and gloo in the CPU tests (world_size 2)."""
from __future__ import annotations

from typing import Tuple


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous slice [rank*n/world, (rank+1)*n/world) of n_total items (pairs, queries or edges)"""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return (rank * n_total) // world, ((rank + 1) * n_total) // world


def is_dist() -> bool:
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized()


def replicate(tensor, src: int = 0):
    """broadcast the tree / obstacle SoA from the owning rank (NVLink/NVSwitch under NCCL); no-op for 1 rank"""
    import torch.distributed as dist
    if is_dist() and dist.get_world_size() > 1:
        dist.broadcast(tensor, src)
    return tensor


def max_over_ranks(value: float, device="cpu") -> float:
    """multi-GPU timings are the max over ranks of the device-measured time"""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if is_dist() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_to_rank0(local, dst: int = 0):
    """results return to the host of rank `dst` (concatenated in rank order); used by tests and examples,
    never inside a timed region"""
    import numpy as np
    import torch.distributed as dist
    if not (is_dist() and dist.get_world_size() > 1):
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    bucket = [None] * world if rank == dst else None
    dist.gather_object(local, bucket, dst=dst)
    return np.concatenate(bucket) if rank == dst else None
