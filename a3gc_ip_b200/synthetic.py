"""Synthetic workloads of BASELINE.json / SURVEY.md section 8d for the bench and the launchers (no dataset is shipped with
the reference): seeded IMU frames of the named shapes and the weight sets of the three-stage pipeline.

    raw frames   acc ~ N(0,1) * std_acc + mean_acc (18-d), ori ~ N(0,1) * std_ori + mean_ori (54-d) per channel, statistics
                 from data/all_sym_train_stats.pt (what evaluate_a3gc_tp.py --norm --cda normalises with)
    prepared x   the same frames after prepare_input: zeros except nodes [3,4,13,14,10] ~ N(0,1)
    weights      stage 1 (H=256): random init, seed 0 (its checkpoint is not shipped); stages 2-3 (H=64, 128): the shipped
                 trained_models/{A3GC,G-GRU} checkpoints (committed as fixtures under tests/golden/weights); AAGC / AGC have no
                 checkpoints at all: random init, seeds 0, 1, 2
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

INPUT_JOINTS = [3, 4, 13, 14, 10]          # evaluate_a3gc_tp.py:65
NUM_NODES = 15
TP_SHAPES = ((12, 3, 256), (15, 3, 64), (15, 9, 128))       # (units_in, units_out, hidden) of stages 1-3 (evaluate_a3gc_tp.py:132-134)
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_GOLDEN = os.path.join(_ROOT, "tests", "golden")


def load_nira() -> Tensor:
    """The 15x15 adjacency template nira_template_15_norm.pkl (evaluate_a3gc_tp.py:128-130), float32."""
    return torch.load(os.path.join(_GOLDEN, "nira_template_15_norm.pt")).float()


def load_stats(sym: bool = True) -> dict:
    """data/all_sym_train_stats.pt (--cda) or data/all_train_stats.pt (evaluate_a3gc_tp.py:66-76)."""
    return torch.load(os.path.join(_GOLDEN, "all_sym_train_stats.pt" if sym else "all_train_stats.pt"))


def synthetic_input(batch: int, steps: int, seed: int) -> Tensor:
    """Prepared stage-1 input [B, T, 15, 12]: zeros except nodes [3,4,13,14,10] ~ N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.zeros(batch, steps, NUM_NODES, 12)
    x[:, :, INPUT_JOINTS, :] = torch.randn(batch, steps, 5, 12, generator=g)
    return x


def synthetic_raw_imu(batch: int, steps: int, seed: int, stats: Optional[dict] = None) -> Tuple[Tensor, Tensor]:
    """Raw frames (ori [B, T, 54], acc [B, T, 18]) whose normalised form is unit normal per channel."""
    g = torch.Generator().manual_seed(seed)
    ori = torch.randn(batch, steps, 54, generator=g)
    acc = torch.randn(batch, steps, 18, generator=g)
    if stats is not None:
        ori = ori * stats["ori"]["std_channel"].float() + stats["ori"]["mean_channel"].float()
        acc = acc * stats["acc"]["std_channel"].float() + stats["acc"]["mean_channel"].float()
    return ori.contiguous(), acc.contiguous()


def _net_class(variant: str):
    from . import net_aagc as N
    return {"AAGC": N.AAGC_net, "A3GC": N.A3GC_net, "AGC": N.AGC_net, "GGRU": N.G_GRU_net}[variant]


def random_state_dict(variant: str, f0: int, out: int, hidden: int, nira: Tensor, seed: int) -> Dict[str, Tensor]:
    """The module's own initialisation (xavier kernels, template adjacencies; net_aagc.py ctors) under a fixed seed, with the
    zero-initialised vectors replaced by small normals so that no term of the cell is trivially absent."""
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        net = _net_class(variant)(f0, out, hidden, nira)
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        for k, v in sd.items():
            if v.dim() == 1:
                sd[k] = 0.1 * torch.randn(v.shape)
    finally:
        torch.random.set_rng_state(gen_state)
    return sd


def trained_state_dict(name: str) -> Dict[str, Tensor]:
    ck = torch.load(os.path.join(_GOLDEN, "weights", name + ".pt"))
    return {k[len("pose_net."):]: v for k, v in ck["state_dict"].items()}


def tp_state_dicts(variant: str, nira: Tensor) -> List[Dict[str, Tensor]]:
    sds = [random_state_dict(variant, *TP_SHAPES[0], nira, seed=0)]
    for i, shape in enumerate(TP_SHAPES[1:], start=2):
        if variant in ("A3GC", "GGRU"):
            sds.append(trained_state_dict(f"{variant}_model{i}"))
        else:
            sds.append(random_state_dict(variant, *shape, nira, seed=i - 1))
    return sds


def build_tp(variant: str, device, engine: str = "auto", precision: str = "fp32", stats: Optional[dict] = None,
             state_dicts: Optional[List[Dict[str, Tensor]]] = None):
    """TPPipeline of three nets with the synthetic-workload weights on `device`; returns (pipeline, state_dicts)."""
    from .pipeline import TPPipeline
    nira = load_nira()
    sds = state_dicts or tp_state_dicts(variant, nira)
    nets = []
    for shape, sd in zip(TP_SHAPES, sds):
        net = _net_class(variant)(*shape, nira)
        net.load_state_dict(sd, strict=True)
        nets.append(net.to(device).eval().set_engine(engine, precision))
    return TPPipeline(*nets, stats=stats), sds
