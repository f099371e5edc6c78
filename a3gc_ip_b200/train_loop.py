"""Training-loop glue around the CUDA training path, following train_a3gc_tp.py:

  * ``stage_inputs``            the per-stage (inputs, target) selection of ``train_one_epoch`` (:57-65): stages are trained
                                independently with teacher-forced inputs, not chained;
  * ``fit_stage``               one stage's epoch loop (:241-262 and its two twins): Adam(lr, weight_decay),
                                ExponentialLR(gamma=0.8) stepped once per epoch, validation in eval mode, checkpoint
                                ``{'epoch': epoch + 1, 'state_dict': ...}`` on every improvement, early stop once
                                ``tolerance_counter > patience``;
  * ``checkpoint_name`` / ``latest_checkpoints``   the file naming ``checkpoint_model{N}_{pretrain|finetuning}_{epoch}.tar``
                                (:255) and the resume discovery by highest epoch number, ``pretrain`` preferred when both
                                kinds are present (:164-187).

  * ``teacher_forced_sample``   what ``GraphDataset_tp.__getitem__`` (datasets.py:40-73) builds per sample, for a batch on the
                                device: normalised + scattered IMU input (the ``prepare_input`` kernel), the noisy teacher
                                inputs ``full_pos + N(0, 0.025)`` (:54) reduced to the leaf / major joints, and the targets.

Data loading and the datasets themselves (datasets.py) stay the reference's.
"""
from __future__ import annotations

import glob
import os
import re
from typing import Callable, Dict, Iterable, Optional, Tuple

import torch

from .sharding import FlatGradAllReducer, train_step


def stage_inputs(model_number: int, imu, leaf_pos_input, full_pos_input, leaf_pos, full_pos, smpl):
    """(inputs, target) of stage 1 / 2 / 3 (train_a3gc_tp.py:57-65)."""
    if model_number == 1:
        return imu, leaf_pos
    if model_number == 2:
        return torch.cat((imu, leaf_pos_input), dim=-1), full_pos
    if model_number == 3:
        return torch.cat((imu, full_pos_input), dim=-1), smpl
    raise ValueError("model_number must be 1, 2 or 3")


LEAF_NODES = [4, 5, 15, 18, 19]                 # datasets.py:21  (SMPL joints)
LEAF_NODES_REDUCED = [3, 4, 10, 13, 14]         # datasets.py:22  (their slots in the 15-node graph)
SMPL_MAJOR_JOINTS = [1, 2, 3, 4, 5, 6, 9, 12, 13, 14, 15, 16, 17, 18, 19]   # datasets.py:23


def teacher_forced_sample(ori: torch.Tensor, acc: torch.Tensor, full_pos: torch.Tensor, smpl: torch.Tensor,
                          stats: Optional[dict] = None, noise_std: float = 0.025, generator: Optional[torch.Generator] = None):
    """Batched, on-device form of ``GraphDataset_tp.__getitem__`` (datasets.py:40-73).

    ori [B,T,54], acc [B,T,18], full_pos [B,T,24,3], smpl [B,T,...] (CUDA) ->
    (inputs [B,T,15,12], leaf_pos_input [B,T,15,3], full_pos_input [B,T,15,3], leaf_pos [B,T,45], full_pos [B,T,45], smpl):
    the tuple ``stage_inputs`` consumes.  The teacher inputs carry the reference's noise ``N(0, noise_std)`` (:54)."""
    from .pipeline import prepare_input
    inputs = prepare_input(ori, acc, stats)
    noise = torch.empty_like(full_pos).normal_(0.0, noise_std, generator=generator)
    full_pos_input = full_pos + noise
    B, T = full_pos.shape[0], full_pos.shape[1]
    leaf_pos = torch.zeros(B, T, 15, 3, dtype=full_pos.dtype, device=full_pos.device)
    leaf_pos_input = torch.zeros_like(leaf_pos)
    leaf_pos[:, :, LEAF_NODES_REDUCED] = full_pos[:, :, LEAF_NODES]
    leaf_pos_input[:, :, LEAF_NODES_REDUCED] = full_pos_input[:, :, LEAF_NODES]
    fp = full_pos[:, :, SMPL_MAJOR_JOINTS]
    fpi = full_pos_input[:, :, SMPL_MAJOR_JOINTS]
    return inputs, leaf_pos_input, fpi.reshape(B, T, 15, 3), leaf_pos.reshape(B, T, 45), fp.reshape(B, T, 45), smpl


def checkpoint_name(model_number: int, epoch: int, finetuning: bool = False) -> str:
    return "checkpoint_model{}_{}_{}.tar".format(model_number, "finetuning" if finetuning else "pretrain", epoch)


def latest_checkpoints(model_path: str) -> Dict[int, str]:
    """Resume discovery of train_a3gc_tp.py:164-187: {1: file, 2: file, 3: file}, each the highest-numbered one."""
    files = glob.glob(os.path.join(model_path, "*"))
    has_pre = any("pretrain" in f for f in files)
    has_fine = any("finetuning" in f for f in files)
    if has_pre and has_fine:
        files = [f for f in files if "pretrain" in f]
    elif not has_pre and not has_fine:
        raise ValueError("Found neither savefiles with pretrain in their name, nor with finetuning")
    out = {}
    for n in (1, 2, 3):
        cand = [(int(re.findall(r"_\d+", os.path.basename(f))[0][1:]), f) for f in files if f"model{n}" in f]
        if cand:
            out[n] = max(cand)[1]
    return out


@torch.no_grad()
def validate(model: torch.nn.Module, criterion, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]], data_parallel: bool = False) -> float:
    """valid_one_epoch (train_a3gc_tp.py:89-125): eval mode (the inference engine), mean loss over the loader.  With
    ``data_parallel`` the sum and the batch count are all-reduced, so every rank returns the SAME global mean."""
    model.eval()
    total, n = 0.0, 0
    for inputs, target in batches:
        prediction, _ = model.forward(inputs, None)
        total += float(criterion.forward(prediction.view(target.shape), target))
        n += 1
    if data_parallel:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dev = next(model.parameters()).device
            acc = torch.tensor([total, float(n)], dtype=torch.float64, device=dev if dist.get_backend() == "nccl" else "cpu")
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)
            total, n = float(acc[0]), int(round(float(acc[1])))
    return total / max(n, 1)


def fit_stage(model: torch.nn.Module, criterion, train_batches: Callable[[], Iterable], valid_batches: Callable[[], Iterable],
              model_number: int, save_dir: Optional[str] = None, lr: float = 1e-3, weight_decay: float = 0.0, patience: int = 3,
              start_epoch: int = 0, max_epochs: int = 500, finetuning: bool = False, data_parallel: bool = False,
              log: Callable[[str], None] = print) -> Dict[str, object]:
    """One stage of train_a3gc_tp.py:241-312.  ``train_batches()`` / ``valid_batches()`` yield (inputs, target) pairs on the
    model's device.  With ``data_parallel`` every rank runs this on its shard (the same NUMBER of batches on every rank --
    each step is a collective): gradients are all-reduced per step weighted by the local batch size, the validation loss is
    the global mean, so the improvement / early-stop decisions are identical on all ranks; rank 0 alone writes the
    checkpoint and the others wait for it."""
    dist = None
    if data_parallel:
        import torch.distributed as _dist
        if _dist.is_available() and _dist.is_initialized() and _dist.get_world_size() > 1:
            dist = _dist
    rank = dist.get_rank() if dist else 0
    optimizer = torch.optim.Adam(model.parameters(), lr, weight_decay=weight_decay)
    scheduler = torch.optim.lr_scheduler.ExponentialLR(optimizer, gamma=0.8)
    for _ in range(start_epoch):                       # the reference does this for stage 3 on resume (:290-292)
        scheduler.step()
    # ragged shards are allowed here, so the all-reduce is issued after backward (no early bucket)
    reducer = FlatGradAllReducer(model.parameters(), overlap=False) if data_parallel else None
    best_loss, tolerance_counter, saved, history = 1e5, 0, None, []
    for epoch in range(start_epoch, max_epochs):
        model.train()
        total, n = 0.0, 0
        for inputs, target in train_batches():
            total += float(train_step(model, criterion, optimizer, inputs, target, reducer))
            n += 1
        scheduler.step()
        train_loss = total / max(n, 1)
        valid_loss = validate(model, criterion, valid_batches(), **({"data_parallel": True} if dist else {}))
        history.append((epoch, train_loss, valid_loss))
        if rank == 0:
            log("|---------- epoch = {}  |  train_loss = {}  |  valid_loss = {} ----------|".format(epoch, train_loss, valid_loss))
        if valid_loss < best_loss:
            tolerance_counter, best_loss = 0, valid_loss
            if save_dir is not None:
                saved = os.path.join(save_dir, checkpoint_name(model_number, epoch, finetuning))
                if rank == 0:
                    torch.save({"epoch": epoch + 1, "state_dict": model.state_dict()}, saved)
                if dist:
                    dist.barrier()
        else:
            tolerance_counter += 1
        if tolerance_counter > patience:
            break
    return {"best_loss": best_loss, "checkpoint": saved, "history": history, "lr": optimizer.param_groups[0]["lr"]}
