"""Host-side sharding of independent camera streams over ranks (one process per GPU).

The reference has no distributed component; its documented scaling recipe is "one tracker per
stream" (docs/guides/architecture.md:249-255).  Streams never interact, so the partition is
static and the data path needs no collective: ranks only agree on the wall-clock window (a
barrier before, max of the elapsed time after) to report an aggregate rate.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def partition_streams(n_streams: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block partition [begin, end) of stream ids for `rank`; sizes differ by at most 1."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_streams, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def owner_of(stream_id: int, n_streams: int, world_size: int) -> int:
    for r in range(world_size):
        b, e = partition_streams(n_streams, world_size, r)
        if b <= stream_id < e:
            return r
    raise ValueError("stream id out of range")


def aggregate_rate(frames_done: int, elapsed_s: float, group=None) -> float:
    """Whole-job frames/s = sum of frames over ranks / max elapsed over ranks.  Uses
    torch.distributed when a process group is up, else the local values."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        t = torch.tensor([float(frames_done)], dtype=torch.float64)
        e = torch.tensor([float(elapsed_s)], dtype=torch.float64)
        if dist.get_backend(group) == "nccl":
            t, e = t.cuda(), e.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(e, op=dist.ReduceOp.MAX, group=group)
        return float(t.item()) / float(e.item())
    return frames_done / elapsed_s


def gather_results(local: Sequence, group=None) -> List:
    """All ranks' per-stream results in stream order (object gather; results are small row lists)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        parts: List = [None] * dist.get_world_size(group)
        dist.all_gather_object(parts, list(local), group=group)
        return [x for p in parts for x in p]
    return list(local)
