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


class ShardedBatch:
    """One ensemble of `n_total` independent land segments split over the ranks of a torch.distributed group (one process per
    GPU): rank r owns the contiguous block shard_bounds(n_total, r, world) -- its parameters, state, forcings and outputs for the
    whole run -- as an ordinary engine.LandBatch on its own device.  No collective touches the data path; `gather()` brings
    per-segment results (reduced outputs, final states, error counts) of all ranks together in segment order (NCCL for CUDA
    tensors, gloo for host tensors).  BASELINE.json configs[2] ("1M-segment PERLND+IMPLND ensemble sharded over 2/4/8 x B200")
    is two of these, one per operation.

    store_of(lo, hi) -> SegmentStore of segments [lo, hi) of the ensemble (e.g. lambda lo, hi: ens.shard(lo, hi).store(cal))."""

    def __init__(self, store_of, n_total, calendar, forcing_names, save, device=0, group=None, **batch_kw):
        from .engine import LandBatch

        rank, _local, ws = world()
        try:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized():
                rank, ws = dist.get_rank(group), dist.get_world_size(group)
        except ImportError:
            pass
        self.rank, self.world, self.group, self.n_total = rank, ws, group, int(n_total)
        self.lo, self.hi = shard_bounds(n_total, rank, ws)
        self.n = self.hi - self.lo
        self.batch = LandBatch(store_of(self.lo, self.hi), calendar, forcing_names, save, device=device, **batch_kw)

    def gather(self, local_columns):
        """[..., n_local] of every rank -> [..., n_total] on every rank"""
        if self.world == 1:
            import torch

            return torch.as_tensor(local_columns)
        return gather_columns(local_columns, self.n_total, self.group)

    def close(self):
        self.batch.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
