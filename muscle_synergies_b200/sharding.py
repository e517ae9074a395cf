"""Partition of independent trials over the GPUs of one box.

One process per GPU, launched with torchrun; every rank loads its own files - the reference
has no cross-file state, so there is NO collective on the data path.  The only exchange is a
host-side gather of small per-file results (status, row counts, VAF tables) through
torch.distributed (gloo or nccl object gather).
"""
from typing import List, Sequence, TypeVar

T = TypeVar("T")


def shard(items: Sequence[T], rank: int, world_size: int, sizes: Sequence[int] = None) -> List[T]:
    """Items of `rank`.  Without sizes: round-robin (item i -> rank i % world).  With sizes
    (bytes per file): greedy longest-processing-time assignment, deterministic on every rank."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    if sizes is None:
        return [it for i, it in enumerate(items) if i % world_size == rank]
    if len(sizes) != len(items):
        raise ValueError("sizes must match items")
    order = sorted(range(len(items)), key=lambda i: (-sizes[i], i))
    load = [0] * world_size
    owner = {}
    for i in order:
        r = min(range(world_size), key=lambda j: (load[j], j))
        owner[i] = r
        load[r] += sizes[i]
    return [items[i] for i in range(len(items)) if owner[i] == rank]


def gather_results(local, group=None):
    """All ranks receive the list of every rank's `local` object (small, host side)."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return [local]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, local, group=group)
    return out
