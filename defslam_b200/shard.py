"""Multi-GPU host logic: independent frames are split into contiguous shards, one per rank; the
only exchange is the reduction of timings / counters ("replicas + one gather")."""
from __future__ import annotations


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous block of ceil(n/world) items per rank (the last ranks may get fewer, or none)"""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per = -(-n_items // world)
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


def reduce_job_stats(local_ms: float, local_solved: int, dist=None, device=None):
    """max over ranks of the time, sum over ranks of the solved frames -> (ms, solved).
    `dist` is torch.distributed (NCCL on the GPU box, gloo in the CPU tests) or None."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(local_ms), int(local_solved)
    import torch
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
    s = torch.tensor([float(local_solved)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return float(t.item()), int(round(s.item()))
