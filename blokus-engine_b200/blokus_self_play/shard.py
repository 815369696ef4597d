"""Sharding of games over the GPUs of one box.

Games are independent (the reference's only parallelism, model/training.py:204-208): rank r owns a
contiguous range of GLOBAL game ids and runs the same kernels on its own GPU; the RNG is keyed by global
id, so results do not depend on the number of ranks.  No collective on the data path — torch.distributed
is used only to gather finished-game tuples on the host and to reduce timings (max) / work (sum)."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """(first_global_id, n_games) of `rank`: contiguous ranges, the remainder spread over the first ranks."""
    if world <= 0 or not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def gather_finished(local: Sequence, dst: int = 0) -> List:
    """Concatenate every rank's finished-game tuples on rank `dst` in global-id order (others get [])."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(local)
    rank, world = dist.get_rank(), dist.get_world_size()
    bucket = [None] * world if rank == dst else None
    dist.gather_object(list(local), bucket, dst=dst)
    if rank != dst:
        return []
    out = []
    for part in bucket:
        out.extend(part)
    return out


def reduce_max_sum(times: Sequence[float], work: Sequence[float], device=None):
    """Timing of a multi-rank run = max over ranks; work = sum over ranks."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(times), dtype=torch.float64, device=device)
    w = torch.tensor(list(work), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    return t.tolist(), w.tolist()
