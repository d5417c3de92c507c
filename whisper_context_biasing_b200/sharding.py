"""Clip sharding for the multi-GPU configurations.

The path shards trivially: every clip is independent (the only cross-element dependency, the max-8
clamp, is inside a clip: TF-FE:156-158).  Rank g of G takes the contiguous clips
[start, stop); one process per GPU, no collective on the data path -- `torch.distributed` is only
used by callers for a start barrier and to combine per-rank timings / checksums.
"""
from __future__ import annotations


def clip_shard(n_clips: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced partition: the first `n_clips % world` ranks get one extra clip."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if n_clips < 0:
        raise ValueError("n_clips must be >= 0")
    base, extra = divmod(n_clips, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_timings(local_ms: float, dist=None, device=None) -> float:
    """Max over ranks of a device-side duration (the multi-GPU number is the slowest rank's)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(local_ms)
    import torch

    t = torch.tensor([local_ms], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
