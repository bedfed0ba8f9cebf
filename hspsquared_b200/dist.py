"""Multi-GPU sharding of a land batch: one process per GPU, contiguous blocks of independent segments, no collective on
the data path.  torch.distributed (NCCL on GPUs, gloo in CPU tests) is used only for the barrier that
brackets timing and for the optional final gather of reduced outputs."""
from __future__ import annotations

import os


def world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_bounds(n_total, rank, world_size):
    """Contiguous, balanced: the first (n_total % world) ranks get one extra segment."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_columns(local, n_total, group=None):
    """All-gather per-segment columns [..., n_local] from every rank into [..., n_total] (rank order = segment order)."""
    import torch
    import torch.distributed as dist

    ws = dist.get_world_size(group)
    t = torch.as_tensor(local)
    sizes = [shard_bounds(n_total, r, ws) for r in range(ws)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(t.shape[:-1] + (width,), dtype=t.dtype, device=t.device)
    pad[..., : t.shape[-1]] = t
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[..., : hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=-1)
