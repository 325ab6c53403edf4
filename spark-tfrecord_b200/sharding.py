"""Multi-GPU plumbing for a path that shards with no exchange step: files (or
record-aligned blocks) are assigned to ranks greedily by size (LPT), every rank decodes its own shard,
and only scalar timing/byte counters are reduced.  The reference's unit of parallelism is the same: one
unsplittable file per Spark task (M/DefaultSource.scala:26-29)."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_lpt(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time-first: indices of `sizes` per rank, deterministic on every rank."""
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    loads = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += sizes[i]
    for r in range(world):
        out[r].sort()
    return out


def aggregate_throughput(local_bytes: float, local_seconds: float, dist=None, device=None) -> Tuple[float, float]:
    """(sum of bytes over ranks, max of seconds over ranks) -- whole-job throughput = bytes / seconds."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(local_bytes), float(local_seconds)
    import torch
    dev = device if device is not None else "cpu"
    b = torch.tensor([float(local_bytes)], dtype=torch.float64, device=dev)
    t = torch.tensor([float(local_seconds)], dtype=torch.float64, device=dev)
    dist.all_reduce(b, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(b.item()), float(t.item())
