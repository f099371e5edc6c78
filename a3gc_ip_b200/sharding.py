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


class FlatGradAllReducer:
    """Data-parallel training glue: one all-reduce(SUM) of a single flat fp32 bucket that
    holds every parameter gradient, then a division by the world size -- the loss is a batch mean
    (net_aagc.py:1086), so the mean of the per-rank gradients is the gradient of the global-batch loss.  The
    largest stage (A3GC, H = 256) has 3.42 M parameters = 13.7 MB: one NCCL call over NVLink per step.

    The bucket is allocated once; `reduce()` copies the grads in (parameters without a gradient contribute
    zeros), all-reduces, and writes the averaged gradients back in place, so any torch optimizer can follow.
    """

    def __init__(self, params, process_group=None):
        import torch.distributed as dist
        self.dist = dist if dist.is_available() and dist.is_initialized() else None
        self.group = process_group
        self.world = self.dist.get_world_size(process_group) if self.dist else 1
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += p.numel()
        self.bucket = torch.zeros(n, dtype=dt, device=dev)
        self.views = [self.bucket[o:o + p.numel()].view_as(p) for o, p in zip(self.offsets, self.params)]

    @property
    def nbytes(self) -> int:
        return self.bucket.numel() * self.bucket.element_size()

    def reduce(self) -> None:
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        if self.dist is not None and self.world > 1:
            self.dist.all_reduce(self.bucket, op=self.dist.ReduceOp.SUM, group=self.group)
            self.bucket.div_(self.world)
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)


def train_step(model: torch.nn.Module, criterion, optimizer, inputs: torch.Tensor, target: torch.Tensor,
               reducer: "FlatGradAllReducer" = None) -> torch.Tensor:
    """One optimisation step exactly as train_a3gc_tp.py:74-84 does it (forward in train mode with rnn_state=None,
    pose loss on the prediction viewed as the target, zero_grad / backward / step), plus the gradient all-reduce
    when data-parallel.  Returns the (local) loss."""
    prediction, _ = model.forward(inputs, None)
    loss = criterion.forward(prediction.view(target.shape), target)
    optimizer.zero_grad()
    loss.backward()
    if reducer is not None:
        reducer.reduce()
    optimizer.step()
    return loss.detach()
