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


def allreduce_schema(local: dict, dist=None, device=None) -> dict:
    """mergeFieldTypes across ranks (the combOp of rdd.aggregate, M/TensorFlowInferSchema.scala:40,43,120-127):
    every rank contributes its name -> lattice-code map; the union dictionary is built deterministically (sorted
    names, all-gathered), then ONE all-reduce(MAX) over the code vector -- MAX is findTightestCommonType because
    null (0) is the identity and otherwise the higher precedence wins (:213-228).  NCCL when `device` is a CUDA
    device (a few KB: latency-bound, one collective), gloo on CPU.  Code 10 (ArrayType(ArrayType(null))) conflicts
    with any other non-null type like in the reference."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(local)
    import torch
    world = dist.get_world_size()
    gathered = [None] * world
    dist.all_gather_object(gathered, sorted(local.keys()))
    names = sorted(set(n for g in gathered for n in g))
    dev = device if device is not None else "cpu"
    codes = torch.tensor([local.get(n, 0) for n in names], dtype=torch.int32, device=dev)
    has10 = (codes == 10).to(torch.int32)
    other = torch.where(codes == 10, torch.zeros_like(codes), codes)
    dist.all_reduce(other, op=dist.ReduceOp.MAX)
    dist.all_reduce(has10, op=dist.ReduceOp.MAX)
    out = {}
    for n, c, h in zip(names, other.tolist(), has10.tolist()):
        if h and c:
            raise RuntimeError("Unable to get the precedence for given datatype (ArrayType(ArrayType(null)) vs another type)")
        out[n] = 10 if h else c
    return out


def codes_to_struct(codes: dict):
    """lattice codes -> StructType (column order is unspecified in the reference: mutable.Map iteration, :48-57;
    here: sorted by name).  Code 0 -> NullType (:50-52)."""
    from .sqltypes import (ArrayType, FloatType, LongType, NullType, StringType, StructField, StructType)
    base = {1: LongType, 2: FloatType, 3: StringType}
    fields = []
    for name in sorted(codes):
        c = codes[name]
        if c == 0:
            t = NullType()
        elif c <= 3:
            t = base[c]()
        elif c <= 6:
            t = ArrayType(base[c - 3]())
        elif c <= 9:
            t = ArrayType(ArrayType(base[c - 6]()))
        else:
            t = ArrayType(ArrayType(NullType()))
        fields.append(StructField(name.decode("utf-8") if isinstance(name, bytes) else name, t))
    return StructType(fields)
