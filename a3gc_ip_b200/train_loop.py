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

Data loading, noise injection and the datasets themselves (datasets.py) stay the reference's.
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
def validate(model: torch.nn.Module, criterion, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]]) -> float:
    """valid_one_epoch (train_a3gc_tp.py:89-125): eval mode (the inference engine), mean loss over the loader."""
    model.eval()
    total, n = 0.0, 0
    for inputs, target in batches:
        prediction, _ = model.forward(inputs, None)
        total += float(criterion.forward(prediction.view(target.shape), target))
        n += 1
    return total / max(n, 1)


def fit_stage(model: torch.nn.Module, criterion, train_batches: Callable[[], Iterable], valid_batches: Callable[[], Iterable],
              model_number: int, save_dir: Optional[str] = None, lr: float = 1e-3, weight_decay: float = 0.0, patience: int = 3,
              start_epoch: int = 0, max_epochs: int = 500, finetuning: bool = False, data_parallel: bool = False,
              log: Callable[[str], None] = print) -> Dict[str, object]:
    """One stage of train_a3gc_tp.py:241-312.  ``train_batches()`` / ``valid_batches()`` yield (inputs, target) pairs on the
    model's device.  With ``data_parallel`` every rank runs this on its shard and gradients are all-reduced per step."""
    optimizer = torch.optim.Adam(model.parameters(), lr, weight_decay=weight_decay)
    scheduler = torch.optim.lr_scheduler.ExponentialLR(optimizer, gamma=0.8)
    for _ in range(start_epoch):                       # the reference does this for stage 3 on resume (:290-292)
        scheduler.step()
    reducer = FlatGradAllReducer(model.parameters()) if data_parallel else None
    best_loss, tolerance_counter, saved, history = 1e5, 0, None, []
    for epoch in range(start_epoch, max_epochs):
        model.train()
        total, n = 0.0, 0
        for inputs, target in train_batches():
            total += float(train_step(model, criterion, optimizer, inputs, target, reducer))
            n += 1
        scheduler.step()
        train_loss = total / max(n, 1)
        valid_loss = validate(model, criterion, valid_batches())
        history.append((epoch, train_loss, valid_loss))
        log("|---------- epoch = {}  |  train_loss = {}  |  valid_loss = {} ----------|".format(epoch, train_loss, valid_loss))
        if valid_loss < best_loss:
            tolerance_counter, best_loss = 0, valid_loss
            if save_dir is not None:
                saved = os.path.join(save_dir, checkpoint_name(model_number, epoch, finetuning))
                torch.save({"epoch": epoch + 1, "state_dict": model.state_dict()}, saved)
        else:
            tolerance_counter += 1
        if tolerance_counter > patience:
            break
    return {"best_loss": best_loss, "checkpoint": saved, "history": history, "lr": optimizer.param_groups[0]["lr"]}
