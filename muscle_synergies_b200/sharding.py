"""Partition of independent trials over the GPUs of one box.

One process per GPU, launched with torchrun; every rank loads its own files - the reference
has no cross-file state, so there is NO collective on the data path.  The only exchange is a
host-side gather of small per-file results (status, row counts, VAF tables) through
torch.distributed (gloo or nccl object gather).
"""
import os
from typing import List, Optional, Sequence, TypeVar

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


def parse_cpulist(text: str) -> List[int]:
    """"0-3,8,10-11" -> [0, 1, 2, 3, 8, 10, 11] (the format of /sys/.../local_cpulist)."""
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(device_index: int) -> Optional[List[int]]:
    """Restricts this process to the CPUs next to GPU `device_index` (its PCIe root's NUMA node), so that
    the pinned staging buffers it allocates afterwards and the threads that fill them are local to the
    GPU - with one process per GPU, host memory traffic is what the ranks compete for (SURVEY.md
    section 8e).  Returns the CPU list, or None when the topology is not exposed (nothing is changed)."""
    import torch

    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            cpus = parse_cpulist(f.read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except (OSError, AttributeError, ValueError):
        return None
