"""CPU oracle for the A3GC-IP recurrent graph-convolution hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: it may be
imported by ``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, and only as the checker or the timed CPU baseline.  The product
path (``a3gc_ip_b200``) never imports it and fails loudly when its CUDA library is missing.

What this is: a plain-PyTorch (CPU, eager) restatement of ``/root/reference/net_aagc.py``
lines 40-695 -- the AAGC graph convolution, the four recurrent cells, the forward / reverse /
bidirectional time loops and the four nets -- written over a flat ``state_dict`` (the
reference's own keys, without the ``pose_net.`` prefix).  Every function cites the reference
lines it follows and keeps the reference's operation order (einsum, then matmul, then bias)
so the fp32 result agrees with the reference to rounding noise.  It runs in fp32 (the
reference's dtype) or fp64 (used as the "truth" when budgeting kernel error).

Pinning: the reference ships no tests or golden vectors for this path,
so the oracle is pinned against *outputs of the reference itself*: ``oracle/gen_golden.py``
imports the unmodified reference from /root/reference in the build container, runs it on
seeded inputs (random-init weights and the shipped ``trained_models`` checkpoints) and commits
the results under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against
them on every CPU test run.

Third-party arithmetic: only PyTorch ATen (mm / bmm / pointwise); the reference pins
pytorch 1.13 in its Dockerfile:1, this container has torch 2.11.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

NUM_NODES = 15                       # net_aagc.py:142, 233, 319 (asserted), :609 (hard-coded)
INPUT_JOINTS = [3, 4, 13, 14, 10]    # evaluate_a3gc_tp.py:65
VARIANTS = ("AAGC", "A3GC", "AGC", "GGRU")

# joint_set.reduced / ignored, config.py:29-30
JOINT_REDUCED = [1, 2, 3, 4, 5, 6, 9, 12, 13, 14, 15, 16, 17, 18, 19]
JOINT_IGNORED = [0, 7, 8, 10, 11, 20, 21, 22, 23]


# --------------------------------------------------------------------------------------
# L0: non-recurrent graph convolution and the four cells
# --------------------------------------------------------------------------------------
def aagc_forward(x: Tensor, sd: StateDict, prefix: str, activation: str = "linear") -> Tensor:
    """``AAGC.forward`` (net_aagc.py:61-66), eval mode (dropout = identity).

    x [B, T, 15, F] -> [B, T, 15, O];  y = act((adj @ x) @ W^T + b).
    """
    adj = sd[prefix + "adj"]
    w = sd[prefix + "gcn_kernel"]
    b = sd[prefix + "gcn_bias"]
    y = torch.einsum("bsnf,nm->bsmf", x, adj.t())          # :63
    y = torch.matmul(y, w.t()) + b                          # :64
    if activation == "tanh":
        y = torch.tanh(y)
    elif activation != "linear":
        raise ValueError("only support linear and tanh activations for now")   # :51
    return y


def _gate(x_s: Tensor, adj: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """One adaptive-adjacency gate pre-activation (net_aagc.py:183-184)."""
    z = torch.einsum("bnf,nm->bmf", x_s, adj.t())
    return torch.matmul(z, w.t()) + b


def _attention(hy: Tensor, sd: StateDict, p: str) -> Tensor:
    """Joint-wise attention block shared by A3GC and AGC cells (net_aagc.py:200-213 / :286-299)."""
    q_t = torch.relu(torch.sum(torch.matmul(hy, sd[p + "attention_w"].t()), dim=1, keepdim=True))
    wh_ht = torch.matmul(hy, sd[p + "attention_wh"].t())
    wq_qt = torch.matmul(q_t, sd[p + "attention_wq"].t())
    qht = torch.tanh(wh_ht + wq_qt + sd[p + "attention_bs"])
    a_t = torch.matmul(qht, sd[p + "attention_u"].t()).squeeze(2) + sd[p + "attention_bu"]
    a_t = torch.sigmoid(a_t.unsqueeze(-1))
    return hy + hy * a_t


def cell_lstm(variant: str, x: Tensor, state: Tuple[Tensor, Tensor], sd: StateDict, p: str,
              activation: str = "tanh") -> Tuple[Tensor, Tuple[Tensor, Tensor]]:
    """AAGC / A3GC / AGC LSTM cells (net_aagc.py:103-126, :178-217, :266-303), eval mode.

    x [B,15,F], state (h, c) [B,15,H] -> (act(h'), (h', c')).  The carried h' has no tanh.
    """
    hx, cx = state
    x_s = torch.cat((x, hx), dim=2)                                             # :182
    if variant == "AGC":
        # one frozen adjacency, applied transposed w.r.t. the A3GC gates (:271)
        x_s = torch.einsum("nm,bmf->bnf", sd[p + "adjacency"].t(), x_s)
        pre = {g: torch.matmul(x_s, sd[p + "gcn_kernel_" + g].t()) + sd[p + "gcn_bias_" + g]
               for g in "ifco"}
    else:
        pre = {g: _gate(x_s, sd[p + "adjacency_" + g], sd[p + "gcn_kernel_" + g],
                        sd[p + "gcn_bias_" + g]) for g in "ifco"}
    x_i = torch.sigmoid(pre["i"])
    x_f = torch.sigmoid(pre["f"])
    x_c = torch.tanh(pre["c"])
    x_o = torch.sigmoid(pre["o"])
    cy = (x_f * cx) + (x_i * x_c)                                               # :196
    hy = x_o * torch.tanh(cy)                                                   # :198
    if variant in ("A3GC", "AGC"):
        hy = _attention(hy, sd, p)
    hyo = torch.tanh(hy) if activation == "tanh" else hy
    return hyo, (hy, cy)


def cell_ggru(x: Tensor, h: Tensor, sd: StateDict, p: str) -> Tuple[Tensor, Tensor]:
    """``G_GRU_cell.forward`` (net_aagc.py:343-368).  Returns (h', h'); activation_fn unused."""
    msg = torch.matmul(h, sd[p + "gcn_kernel"].t())                             # :347
    msg = torch.einsum("nm,bmf->bnf", sd[p + "adjacency"].t(), msg)             # :348

    def lin(name: str, v: Tensor, bias: bool) -> Tensor:
        y = torch.matmul(v, sd[p + name + ".weight"].t())
        return y + sd[p + name + ".bias"] if bias else y

    r = torch.sigmoid(lin("dense_r_in", x, True) + lin("dense_r_hid", msg, False))
    u = torch.sigmoid(lin("dense_u_in", x, True) + lin("dense_u_hid", msg, False))
    c = torch.tanh(lin("dense_c_in", x, True) + r * lin("dense_c_hid", msg, False))
    h = u * h + (1 - u) * c                                                     # :364
    return h, h


# --------------------------------------------------------------------------------------
# L1: time loops
# --------------------------------------------------------------------------------------
def layer_forward(variant: str, x_tb: Tensor, state, sd: StateDict, p: str, reverse: bool):
    """Forward / Reverse layer (net_aagc.py:435-441, :449-456 and the AAGC/AGC/G_GRU twins).

    x_tb is time-major [T, B, 15, F].  The reverse layer walks t = T-1..0 and returns its
    outputs re-ordered to ascending t; its final state is the state after t = 0.
    """
    T = x_tb.shape[0]
    outs: List[Optional[Tensor]] = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        if variant == "GGRU":
            o, state = cell_ggru(x_tb[t], state, sd, p + "cell.")
        else:
            o, state = cell_lstm(variant, x_tb[t], state, sd, p + "cell.")
        outs[t] = o
    return torch.stack(outs), state


def bi_layer_forward(variant: str, x: Tensor, states: Sequence, sd: StateDict, p: str):
    """``Bi*`` layers (net_aagc.py:469-480): batch-major in/out, directions run on states[i]."""
    x_tb = torch.transpose(x, 0, 1)
    outs, out_states = [], []
    for i in range(2):
        o, s = layer_forward(variant, x_tb, states[i], sd, f"{p}directions.{i}.", reverse=(i == 1))
        outs.append(torch.transpose(o, 0, 1))
        out_states.append(s)
    return torch.cat(outs, -1), out_states


# --------------------------------------------------------------------------------------
# L2: nets
# --------------------------------------------------------------------------------------
def net_forward(variant: str, x: Tensor, sd: StateDict, h=None, prefix: str = ""):
    """``AAGC_net / A3GC_net / AGC_net / G_GRU_net .forward`` (net_aagc.py:607-619, :633-645,
    :659-671, :685-695), eval mode.  x [B,T,15,F0] -> (y [B,T,15,O], rnn2 final states).

    rnn2 is seeded with rnn1's final states (:642-643).
    """
    if variant not in VARIANTS:
        raise ValueError(variant)
    hidden = sd[prefix + "linear_in.gcn_kernel"].shape[0]
    if h is None:
        z = torch.zeros(x.size(0), NUM_NODES, hidden, dtype=x.dtype, device=x.device)
        h = [z, z.clone()] if variant == "GGRU" else [(z, z.clone()), (z.clone(), z.clone())]
    y = aagc_forward(x, sd, prefix + "linear_in.")
    y = torch.relu(y)
    y, h = bi_layer_forward(variant, y, h, sd, prefix + "rnn1.")
    y, h = bi_layer_forward(variant, y, h, sd, prefix + "rnn2.")
    y = aagc_forward(y, sd, prefix + "linear_out.")
    return y, h


def tp_forward(variant: str, x: Tensor, sds: Sequence[StateDict], prefix: str = "") -> Tuple[Tensor, Tensor, Tensor]:
    """Three-stage "TP" chaining of evaluate_a3gc_tp.py:164-172, generalised from B=1 to B.

    x [B,T,15,12] -> (leaf_pos [B,T,15,3], full_pos [B,T,15,3], pose [B,T,15,9]).
    """
    y1, _ = net_forward(variant, x, sds[0], prefix=prefix)
    y2, _ = net_forward(variant, torch.cat((x, y1), dim=-1), sds[1], prefix=prefix)    # :168-169
    y3, _ = net_forward(variant, torch.cat((x, y2), dim=-1), sds[2], prefix=prefix)    # :170-171
    return y1, y2, y3


def pose_loss(pred: Tensor, targ: Tensor, loss_weight: Optional[Tensor] = None) -> Tensor:
    """``pose_loss.forward`` (net_aagc.py:1081-1087): sum of squared error over the last dim, mean over the rest."""
    l = torch.square(targ - pred)
    if loss_weight is not None:
        l = l * loss_weight
    return torch.mean(torch.sum(l, -1, keepdim=False))


# --------------------------------------------------------------------------------------
# caller-side data formats (evaluate_a3gc_tp.py)
# --------------------------------------------------------------------------------------
def prepare_input(ori: Tensor, acc: Tensor, stats: Optional[dict] = None) -> Tensor:
    """``prepare_input`` (evaluate_a3gc_tp.py:64-94) for one recording or a batch.

    ori [..., T, 54] (6 IMUs x 3x3), acc [..., T, 18] (6 x 3) -> [..., T, 15, 12]: normalise per
    channel (``stats`` = the loaded ``data/all*_train_stats.pt`` dict; None = --norm off), drop the
    6th (root) IMU, concatenate (acc, ori) per IMU and scatter onto nodes [3, 4, 13, 14, 10].
    """
    ori = ori.float()
    acc = acc.float()
    if stats is not None:
        ori = (ori - stats["ori"]["mean_channel"]) / stats["ori"]["std_channel"]     # :78
        acc = (acc - stats["acc"]["mean_channel"]) / stats["acc"]["std_channel"]     # :79
    lead = ori.shape[:-1]
    inputs_ = torch.cat((acc.reshape(*lead, 6, 3)[..., :5, :], ori.reshape(*lead, 6, 9)[..., :5, :]), dim=-1)  # :90
    out = torch.zeros(*lead, NUM_NODES, 12, dtype=inputs_.dtype)
    for i, el in enumerate(INPUT_JOINTS):                                            # :91-92
        out[..., el, :] = inputs_[..., i, :]
    return out


def reduced_to_full(reduced_pose: Tensor) -> Tensor:
    """``reduced_to_full`` (evaluate_a3gc_tp.py:59-62): scatter 15 reduced joints into 24, identity elsewhere."""
    full = torch.eye(3, dtype=reduced_pose.dtype).repeat(reduced_pose.shape[0], 24, 1, 1)
    full[:, JOINT_REDUCED] = reduced_pose
    return full


# SMPL 24-joint kinematic tree (kintree_table[0] of the SMPL model file, articulate/model.py:37; the file itself is not
# shipped with the reference -- config.py:23 -- this is the published SMPL topology, also recorded in tests/golden/ik_cases.pt)
SMPL_PARENT = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]


def r6d_to_rotation_matrix(r6d: Tensor) -> Tensor:
    """``articulate.math.r6d_to_rotation_matrix`` (articulate/math/angular.py:167-182): Gram-Schmidt of two 3-vectors."""
    r6d = r6d.reshape(-1, 6)
    nrm = lambda v: v / v.norm(dim=1, keepdim=True)
    c0 = nrm(r6d[:, 0:3])
    c1 = nrm(r6d[:, 3:6] - (c0 * r6d[:, 3:6]).sum(dim=1, keepdim=True) * c0)
    c2 = torch.cross(c0, c1, dim=1)
    r = torch.stack((c0, c1, c2), dim=-1)
    r[torch.isnan(r)] = 0
    return r


def reduced_global_to_full_local(glb_reduced_pose: Tensor, rotsize: int = 9) -> Tensor:
    """``PoseNet3._reduced_glb_to_full_local_mat`` / ``_reduced_glb_6d_to_full_local_mat`` (net_aagc.py:788-800):
    [N, 15, 3, 3] (or [N, 15, 6]) -> [N, 24, 3, 3]; scatter by joint_set.reduced, inverse kinematics along the SMPL
    tree (R_local[i] = R_global[parent[i]]^T R_global[i], articulate/math/spatial.py:115-123, 197-221), identity on
    joint_set.ignored."""
    if rotsize == 6:
        glb_reduced_pose = r6d_to_rotation_matrix(glb_reduced_pose).view(-1, NUM_NODES, 3, 3)
    g = reduced_to_full(glb_reduced_pose.reshape(-1, NUM_NODES, 3, 3))
    local = [g[:, 0]]
    for i in range(1, 24):
        local.append(torch.bmm(g[:, SMPL_PARENT[i]].transpose(1, 2), g[:, i]))
    pose = torch.stack(local, dim=1)
    pose[:, JOINT_IGNORED] = torch.eye(3, dtype=pose.dtype)
    return pose


# --------------------------------------------------------------------------------------
# parameter tables and seeded random weights
# --------------------------------------------------------------------------------------
def cell_param_shapes(variant: str, f_in: int, hidden: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """Registration-order (= state_dict order) parameter table of each cell."""
    H, F, N = hidden, f_in, NUM_NODES
    att = [("attention_w", (H, H)), ("attention_wq", (H, H)), ("attention_wh", (H, H)),
           ("attention_u", (1, H)), ("attention_bs", (H,)), ("attention_bu", (N,))]
    kern = [(f"gcn_kernel_{g}", (H, F + H)) for g in "ifco"]
    adjs = [(f"adjacency_{g}", (N, N)) for g in "ifco"]
    bias = [(f"gcn_bias_{g}", (H,)) for g in "ifco"]
    if variant == "AAGC":                                  # net_aagc.py:84-95
        return kern + adjs + bias
    if variant == "A3GC":                                  # :147-165
        return kern + adjs + bias + att
    if variant == "AGC":                                   # :238-253
        return [("adjacency", (N, N))] + kern + bias + att
    if variant == "GGRU":                                  # :324-335 (module registration order)
        out = [("a", (N, N)), ("adjacency", (N, N)), ("gcn_kernel", (H, H))]
        for g in "ruc":
            out += [(f"dense_{g}_in.weight", (H, F)), (f"dense_{g}_in.bias", (H,))]
        for g in "ruc":
            out += [(f"dense_{g}_hid.weight", (H, H))]
        return out
    raise ValueError(variant)


def net_param_shapes(variant: str, f0: int, out: int, hidden: int) -> List[Tuple[str, Tuple[int, ...]]]:
    H = hidden
    tbl = [("linear_in.gcn_kernel", (H, f0)), ("linear_in.adj", (NUM_NODES, NUM_NODES)), ("linear_in.gcn_bias", (H,))]
    for layer, f_in in (("rnn1", H), ("rnn2", 2 * H)):
        for d in range(2):
            tbl += [(f"{layer}.directions.{d}.cell.{k}", s) for k, s in cell_param_shapes(variant, f_in, H)]
    tbl += [("linear_out.gcn_kernel", (out, 2 * H)), ("linear_out.adj", (NUM_NODES, NUM_NODES)), ("linear_out.gcn_bias", (out,))]
    return tbl


def random_state_dict(variant: str, f0: int, out: int, hidden: int, adjacency: Tensor, seed: int,
                      dtype=torch.float32) -> StateDict:
    """Seeded weights with EVERY parameter randomised (biases, attention_bs/bu and adjacency noise
    included) so that indexing / transposition bugs cannot hide behind zero-init biases or
    identical adjacencies.  Magnitudes follow the reference's init
    (xavier_uniform kernels, template-transposed adjacencies) with small perturbations.
    """
    g = torch.Generator().manual_seed(seed)
    sd: StateDict = {}
    for name, shape in net_param_shapes(variant, f0, out, hidden):
        leaf = name.split(".")[-1]
        if len(shape) == 2 and shape == (NUM_NODES, NUM_NODES):
            t = adjacency.t().to(torch.float32) + 0.05 * torch.randn(shape, generator=g)
        elif len(shape) == 2:
            fan_out, fan_in = shape
            bound = (6.0 / (fan_in + fan_out)) ** 0.5
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:
            t = 0.1 * torch.randn(shape, generator=g)
        sd[name] = t.to(dtype).contiguous()
        del leaf
    return sd


def synthetic_input(batch: int, steps: int, seed: int, f0_extra: int = 0) -> Tensor:
    """Synthetic stage-1 input of SURVEY.md section 8d: zeros except nodes [3,4,13,14,10] ~ N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.zeros(batch, steps, NUM_NODES, 12 + f0_extra)
    x[:, :, INPUT_JOINTS, :12] = torch.randn(batch, steps, 5, 12, generator=g)
    return x


def cast_sd(sd: StateDict, dtype) -> StateDict:
    return {k: v.to(dtype) for k, v in sd.items()}
