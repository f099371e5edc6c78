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
    """Data-parallel training glue: the gradients of all parameters live in ONE flat fp32 bucket
    (every ``p.grad`` is a view into it, so autograd accumulates straight into the bucket and nothing is copied in or out),
    all-reduced (SUM) over NCCL / NVLink and divided by the global batch weight -- the loss is a batch mean
    (net_aagc.py:1086), so the batch-size-weighted mean of the per-rank gradients is the gradient of the global-batch
    loss.  The largest stage (A3GC, H = 256) has 3.42 M parameters = 13.7 MB.

    Overlap: the bucket is cut in two at ``split`` (parameters in registration order).  Backward finalises the gradients in
    reverse order (linear_out, rnn2, then rnn1, linear_in), so the BACK part [split:] is complete while rnn1 is still
    walking its reverse-time chain: a post-accumulate hook issues its all-reduce on a side stream at that moment; the FRONT
    part follows in ``reduce()`` after backward.  ``overlap = False`` issues both after backward (the serial form).

    Use: ``red.zero_grad()`` instead of ``optimizer.zero_grad()`` (keeps the views), ``loss.backward()``,
    ``red.reduce(local_batch)``, ``optimizer.step()``  -- see ``train_step``.
    """

    def __init__(self, params, process_group=None, split: int = 0, overlap: bool = True):
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
        self.split = min(max(int(split), 0), len(self.params))
        self.split_off = self.offsets[self.split] if self.split < len(self.params) else n
        self.overlap = overlap
        self._pending = 0
        self._back_issued = False
        self._equal_checked = False
        self._work = []
        self._comm = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._hooks = []
        if self.split > 0 and self.dist is not None and self.world > 1:
            for p in self.params[self.split:]:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    @classmethod
    def for_net(cls, net: torch.nn.Module, process_group=None, overlap: bool = True) -> "FlatGradAllReducer":
        """Reducer of one of the nets: the bucket is cut between rnn1 and rnn2 (linear_in + rnn1 | rnn2 + linear_out)."""
        named = [(n, p) for n, p in net.named_parameters() if p.requires_grad]
        split = sum(1 for n, _ in named if n.startswith(("linear_in.", "rnn1.", "pose_net.linear_in.", "pose_net.rnn1.")))
        return cls([p for _, p in named], process_group, split=split, overlap=overlap)

    def close(self) -> None:
        """Detach the post-accumulate hooks (a reducer that is replaced by another one on the same parameters)."""
        for h in self._hooks:
            h.remove()
        self._hooks = []

    @property
    def nbytes(self) -> int:
        return self.bucket.numel() * self.bucket.element_size()

    def zero_grad(self) -> None:
        """Zero the bucket and (re)bind every ``p.grad`` to its view of it."""
        self.bucket.zero_()
        for v, p in zip(self.views, self.params):
            if p.grad is not v:
                p.grad = v
        self._pending = len(self.params) - self.split
        self._back_issued = False
        self._work = []

    def _all_reduce(self, lo: int, hi: int) -> None:
        if hi <= lo:
            return
        part = self.bucket[lo:hi]
        if self._comm is not None:
            self._comm.wait_stream(torch.cuda.current_stream(part.device))   # the gradients written so far on the compute stream
            with torch.cuda.stream(self._comm):
                self.dist.all_reduce(part, op=self.dist.ReduceOp.SUM, group=self.group)
        else:
            self.dist.all_reduce(part, op=self.dist.ReduceOp.SUM, group=self.group)

    def _on_grad(self, p) -> None:
        self._pending -= 1
        if self._pending == 0 and self.overlap and not self._back_issued:
            self._back_issued = True
            self._all_reduce(self.split_off, self.bucket.numel())

    def reduce(self, local_batch: int = 1) -> None:
        """After backward: finish the all-reduce and turn the sums into the global-batch mean.  ``local_batch`` = sequences
        this rank contributed (ranks may hold unequal, even empty, shards): every rank's gradient of ITS batch mean is
        weighted by its share of the global batch."""
        for v, p in zip(self.views, self.params):                 # a gradient that was re-created outside the bucket
            if p.grad is None:
                v.zero_()
                p.grad = v
            elif p.grad is not v:
                v.copy_(p.grad)
                p.grad = v
        if self.dist is None or self.world == 1:
            return
        # total batch weight on the device (no host synchronisation: the stage steps of the three nets stay concurrent)
        tot = torch.full((1,), float(local_batch), dtype=self.bucket.dtype, device=self.bucket.device)
        if self._back_issued:
            self.bucket[:self.split_off].mul_(float(local_batch))
            self._all_reduce(0, self.split_off)
            # the back part went out unweighted: exact only when all shards are equal -- verified once, below
        else:
            self.bucket.mul_(float(local_batch))
            self._all_reduce(0, self.bucket.numel())
        if self._comm is not None:
            self._comm.wait_stream(torch.cuda.current_stream(self.bucket.device))
            with torch.cuda.stream(self._comm):
                self.dist.all_reduce(tot, op=self.dist.ReduceOp.SUM, group=self.group)
            torch.cuda.current_stream(self.bucket.device).wait_stream(self._comm)
        else:
            self.dist.all_reduce(tot, op=self.dist.ReduceOp.SUM, group=self.group)
        if self._back_issued:
            if not self._equal_checked:                           # one host read, on the first overlapped step only
                self._equal_checked = True
                total = float(tot.item())
                if abs(total - self.world * float(local_batch)) > 1e-6 * max(total, 1.0):
                    raise RuntimeError("FlatGradAllReducer: overlap=True needs equal per-rank batches (use overlap=False for ragged shards)")
            self.bucket[self.split_off:].mul_(float(local_batch))
        self.bucket.div_(tot.clamp_min(1e-30))


def train_step(model: torch.nn.Module, criterion, optimizer, inputs: torch.Tensor, target: torch.Tensor,
               reducer: "FlatGradAllReducer" = None) -> torch.Tensor:
    """One optimisation step exactly as train_a3gc_tp.py:74-84 does it (forward in train mode with rnn_state=None,
    pose loss on the prediction viewed as the target, zero_grad / backward / step), plus the gradient all-reduce
    when data-parallel.  Returns the (local) loss."""
    prediction, _ = model.forward(inputs, None)
    loss = criterion.forward(prediction.view(target.shape), target)
    if reducer is not None:
        reducer.zero_grad()
    else:
        optimizer.zero_grad()
    loss.backward()
    if reducer is not None:
        reducer.reduce(inputs.shape[0])
    optimizer.step()
    return loss.detach()
