"""Host-side sharding of independent pairs / sequences over ranks (one process per GPU).

The UAHN forward of a pair depends only on its two frames and its prior, so shards never exchange data:
there is NO collective on the timed path.  torch.distributed is used for the start/stop
barrier and for reducing the device time to its max over ranks.
"""
from __future__ import annotations


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of `n_items` sequences for `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_max_ms(ms: float, device=None) -> float:
    """Max over ranks of a per-rank device time (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms)
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
