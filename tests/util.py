"""Shared helpers for the parity tests."""
import os

import torch

from conftest import load_golden
from oracle import net_oracle as O

NET_CLS_NAMES = {"AAGC": "AAGC_net", "A3GC": "A3GC_net", "AGC": "AGC_net", "GGRU": "G_GRU_net"}
CELL_CLS_NAMES = {"AAGC": "AAGC_LSTM_cell", "A3GC": "A3GC_LSTM_cell", "AGC": "AGC_LSTM_cell", "GGRU": "G_GRU_cell"}


def case_sd(case, nira):
    if "sd" in case:
        return case["sd"]
    if "weights" in case:
        return trained_sd(case["weights"])
    return O.random_state_dict(case["variant"], case["f0"], case["out"], case["hidden"], nira, seed=case["sd_seed"])


def trained_sd(name):
    ck = load_golden(os.path.join("weights", name + ".pt"))
    return {k[len("pose_net."):]: v for k, v in ck["state_dict"].items()}


def unflatten_h(variant, flat, device=None):
    if flat is None:
        return None
    f = [t.to(device) if device is not None else t for t in flat]
    if variant == "GGRU":
        return [f[0], f[1]]
    return [(f[0], f[1]), (f[2], f[3])]


def flatten_h(h):
    out = []
    for s in h:
        out += list(s) if isinstance(s, (tuple, list)) else [s]
    return out


def build_net(variant, f0, out, hidden, sd, nira, device="cuda", engine="auto", precision="fp32"):
    import a3gc_ip_b200 as A
    net = getattr(A, NET_CLS_NAMES[variant])(f0, out, hidden, nira.float())
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return net.to(device).eval().set_engine(engine, precision)


def tp_state_dicts(variant, nira):
    """Weights of BASELINE cfg 1/2: stage 1 random-init seed 0, stages 2-3 trained where shipped."""
    sd1 = O.random_state_dict(variant, 12, 3, 256, nira, seed=0)
    if variant in ("A3GC", "GGRU"):
        return [sd1, trained_sd(f"{variant}_model2"), trained_sd(f"{variant}_model3")]
    return [sd1, O.random_state_dict(variant, 15, 3, 64, nira, seed=1), O.random_state_dict(variant, 15, 9, 128, nira, seed=2)]


def build_tp(variant, nira, device="cuda", engine="auto", precision="fp32"):
    import a3gc_ip_b200 as A
    sds = tp_state_dicts(variant, nira)
    shapes = ((12, 3, 256), (15, 3, 64), (15, 9, 128))
    nets = [build_net(variant, f0, o, h, sd, nira, device, engine, precision) for (f0, o, h), sd in zip(shapes, sds)]
    return A.TPPipeline(*nets), sds
