"""Batch-sharding launcher: independent sequences shard by batch across the GPUs of a box with no
inter-GPU traffic in inference.  One process per GPU; rank r of W owns a
contiguous slice of the batch.  Weights are replicated (<= 13.7 MB per stage)."""
from __future__ import annotations

import os
from typing import Callable, Tuple

import torch


def shard_range(batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of rank `rank`; sizes differ by at most one, empty shards allowed."""
    if world_size <= 0 or not (0 <= rank < world_size) or batch < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedRunner:
    """Runs `fn(x_shard)` on this rank's slice of a host batch and (optionally) gathers the results.

    With torch.distributed initialised (torchrun) it uses RANK / WORLD_SIZE; the data path needs no
    collective -- `gather` exists for callers that want the full output on every rank (gloo or nccl).
    """

    def __init__(self, device: torch.device = None):
        import torch.distributed as dist
        self.dist = dist if dist.is_available() and dist.is_initialized() else None
        self.rank = self.dist.get_rank() if self.dist else int(os.environ.get("RANK", 0))
        self.world = self.dist.get_world_size() if self.dist else int(os.environ.get("WORLD_SIZE", 1))
        self.device = device

    def shard(self, x: torch.Tensor) -> torch.Tensor:
        lo, hi = shard_range(x.shape[0], self.world, self.rank)
        return x[lo:hi]

    def run(self, fn: Callable[[torch.Tensor], torch.Tensor], x: torch.Tensor) -> torch.Tensor:
        xs = self.shard(x)
        if self.device is not None:
            xs = xs.to(self.device, non_blocking=True)
        return fn(xs)

    def gather(self, y_local: torch.Tensor, batch: int) -> torch.Tensor:
        """All ranks receive the concatenated [batch, ...] result (ragged shards handled by padding)."""
        if self.dist is None or self.world == 1:
            return y_local
        sizes = [shard_range(batch, self.world, r) for r in range(self.world)]
        mx = max(hi - lo for lo, hi in sizes)
        pad = torch.zeros((mx,) + tuple(y_local.shape[1:]), dtype=y_local.dtype, device=y_local.device)
        pad[: y_local.shape[0]] = y_local
        outs = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(outs, pad)
        return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)
