"""Training path (BPTT) of the LSTM-family layers: the autograd glue around ``a3gc_layer_train_forward`` and
``a3gc_layer_backward`` (include/a3gc_b200.h).

The reference trains by calling ``model.forward(inputs, rnn_state=None)`` in ``train()`` mode and letting
autograd unroll the TorchScript time loop (train_a3gc_tp.py:74-84).  Here the forward keeps a tape, the
recurrent gradient chain runs in one CUDA kernel per layer (reverse time, both directions concurrently), and
what is left after the chain -- the weight / adjacency / input gradients, which are sums over all (t, b) of
outer products of per-step tensors the chain has stored -- are plain batched GEMMs, issued through torch
(cuBLAS) on the caller's stream.  The large ones run on the tensor cores as three TF32 passes over operands split
into an exactly-TF32 head and an fp32 remainder (``a3gc_train_split_tf32``; hi*hi + lo*hi + hi*lo, error 2^-22 per
product); ``A3GC_TRAIN_GEMM=fp32`` selects plain fp32 SGEMM instead.

Dropout (net_aagc.py:180-181): input dropout is applied to x by the caller of the layer; recurrent dropout is a
Bernoulli mask drawn here with torch's generator and applied inside both kernels.  The reference's masks come
from the TorchScript interpreter's RNG calls and cannot be reproduced bit for bit; gradient parity is tested
with dropout = 0.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib

NUM_NODES = 15
LSTM_PARAM_NAMES = {
    "AAGC": [f"gcn_kernel_{g}" for g in "ifco"] + [f"adjacency_{g}" for g in "ifco"] + [f"gcn_bias_{g}" for g in "ifco"],
    "A3GC": [f"gcn_kernel_{g}" for g in "ifco"] + [f"adjacency_{g}" for g in "ifco"] + [f"gcn_bias_{g}" for g in "ifco"]
            + ["attention_w", "attention_wq", "attention_wh", "attention_u", "attention_bs", "attention_bu"],
    "AGC": [f"gcn_kernel_{g}" for g in "ifco"] + ["adjacency"] + [f"gcn_bias_{g}" for g in "ifco"]
           + ["attention_w", "attention_wq", "attention_wh", "attention_u", "attention_bs", "attention_bu"],
}


# ----------------------------------------------------------------------------------------------------------------------
# hoisted GEMMs: fp32 accuracy from three TF32 tensor-core passes
# ----------------------------------------------------------------------------------------------------------------------
def _gemm_mode() -> str:
    mode = os.environ.get("A3GC_TRAIN_GEMM", "mixed")
    if mode not in ("mixed", "tf32x3", "fp32"):
        raise ValueError(f"A3GC_TRAIN_GEMM must be mixed, tf32x3 or fp32, got {mode!r}")
    return mode


class _Split:
    """An fp32 matrix as (hi, lo): hi exactly representable in TF32, lo = x - hi.  In fp32 mode hi = x, lo = None.
    Mixed mode: hi (fp32) plus bf16 copies hi16 = bf16(hi), lo16 = bf16(lo) for the two correction products (lo = None)."""
    __slots__ = ("hi", "lo", "hi16", "lo16")

    def __init__(self, hi: Tensor, lo: Optional[Tensor], hi16: Optional[Tensor] = None, lo16: Optional[Tensor] = None):
        self.hi, self.lo, self.hi16, self.lo16 = hi, lo, hi16, lo16

    def cols(self, a: int, b: int) -> "_Split":
        cut = lambda t: None if t is None else t[:, a:b]
        return _Split(self.hi[:, a:b], cut(self.lo), cut(self.hi16), cut(self.lo16))


def _mixed_buffers(rows: int, ld: int, device) -> _Split:
    return _Split(torch.empty(rows, ld, dtype=torch.float32, device=device), None,
                  torch.empty(rows, ld, dtype=torch.bfloat16, device=device), torch.empty(rows, ld, dtype=torch.bfloat16, device=device))


def _split_into(dst: _Split, x: Tensor, col0: int) -> None:
    """Mixed split of x [rows, cols] into columns col0.. of the row-major buffers of dst (a3gc_train_split_mixed)."""
    x = x.contiguous()
    rows, cols = x.shape
    with torch.cuda.device(x.device):
        rc = _lib.lib().a3gc_train_split_mixed(x.data_ptr(), rows, cols, dst.hi.data_ptr(), dst.hi16.data_ptr(), dst.lo16.data_ptr(),
                                               dst.hi.shape[1], col0, _lib.stream_ptr(x.device))
    _lib.check(rc, "a3gc_train_split_mixed")


def _split(x: Tensor) -> _Split:
    x = x.contiguous()
    mode = _gemm_mode()
    if mode == "fp32":
        return _Split(x, None)
    if mode == "mixed" and x.dim() == 2 and x.shape[1] % 4 == 0:
        out = _mixed_buffers(x.shape[0], x.shape[1], x.device)
        _split_into(out, x, 0)
        return out
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = _lib.lib().a3gc_train_split_tf32(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), _lib.stream_ptr(x.device))
    _lib.check(rc, "a3gc_train_split_tf32")
    return _Split(hi, lo)


def _hprev_split(hp: Tensor, h0: Optional[Tensor], mask: Optional[Tensor], reverse: int, into: Optional[_Split] = None,
                 col0: int = 0) -> _Split:
    """Rows of the h half of S = [x | h_prev] of one direction, [B*T*15, H] (see a3gc_train_hprev_split); with `into`
    (mixed mode) they are written into columns col0.. of an existing S buffer instead."""
    B, T, _, H = hp.shape
    if into is not None:
        with torch.cuda.device(hp.device):
            rc = _lib.lib().a3gc_train_hprev_split_mixed(hp.data_ptr(), _lib.ptr(h0), _lib.ptr(mask), into.hi.data_ptr(),
                                                         into.hi16.data_ptr(), into.lo16.data_ptr(), B, T, H, into.hi.shape[1], col0,
                                                         int(reverse), _lib.stream_ptr(hp.device))
        _lib.check(rc, "a3gc_train_hprev_split_mixed")
        return into
    if _gemm_mode() == "fp32" or H % 4 != 0:
        hprev = torch.cat((hp[:, 1:], h0.unsqueeze(1)), dim=1) if reverse else torch.cat((h0.unsqueeze(1), hp[:, :-1]), dim=1)
        if mask is not None:
            hprev = hprev * mask
        return _split(hprev.reshape(B * T * NUM_NODES, H))
    if _gemm_mode() == "mixed":
        return _hprev_split(hp, h0, mask, reverse, _mixed_buffers(B * T * NUM_NODES, H, hp.device), 0)
    hi, lo = torch.empty(B * T * NUM_NODES, H, dtype=torch.float32, device=hp.device), torch.empty(B * T * NUM_NODES, H, dtype=torch.float32, device=hp.device)
    with torch.cuda.device(hp.device):
        rc = _lib.lib().a3gc_train_hprev_split(hp.data_ptr(), _lib.ptr(h0), _lib.ptr(mask), hi.data_ptr(), lo.data_ptr(),
                                               B, T, H, int(reverse), _lib.stream_ptr(hp.device))
    _lib.check(rc, "a3gc_train_hprev_split")
    return _Split(hi, lo)


class _tf32_passes:
    """Scope in which torch's matmuls may use TF32 tensor-core kernels (operands are pre-split, so nothing is lost)."""

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.fp32_precision
        torch.backends.cuda.matmul.fp32_precision = "tf32"

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.fp32_precision = self.prev
        return False


def _pair_forms(a: _Split, b: _Split) -> None:
    """A mixed operand meeting a three-pass one (a width that is not a multiple of 4): give the latter bf16 copies too."""
    for s_, o_ in ((a, b), (b, a)):
        if s_.hi16 is None and s_.lo is not None and o_.hi16 is not None:
            s_.hi16, s_.lo16 = s_.hi.bfloat16(), s_.lo.bfloat16()


def _mm_tn(a: _Split, b: _Split) -> Tensor:
    """a^T @ b for row-aligned a [R, M], b [R, N]."""
    _pair_forms(a, b)
    if a.hi16 is not None and b.hi16 is not None:
        out = torch.mm(a.lo16.t(), b.hi16, out_dtype=torch.float32)
        out += torch.mm(a.hi16.t(), b.lo16, out_dtype=torch.float32)
        with _tf32_passes():
            out.addmm_(a.hi.t(), b.hi)
        return out
    if a.lo is None:
        return a.hi.t() @ b.hi
    with _tf32_passes():
        out = a.lo.t() @ b.hi
        out.addmm_(a.hi.t(), b.lo)
        out.addmm_(a.hi.t(), b.hi)
    return out


def _addmm_nn(out: Tensor, a: _Split, b: _Split) -> None:
    """out += a @ b for a [R, K], b [K, N]."""
    _pair_forms(a, b)
    if a.hi16 is not None and b.hi16 is not None:
        out += torch.mm(a.lo16, b.hi16, out_dtype=torch.float32)
        out += torch.mm(a.hi16, b.lo16, out_dtype=torch.float32)
        with _tf32_passes():
            out.addmm_(a.hi, b.hi)
        return
    if a.lo is None:
        out.addmm_(a.hi, b.hi)
        return
    with _tf32_passes():
        out.addmm_(a.lo, b.hi)
        out.addmm_(a.hi, b.lo)
        out.addmm_(a.hi, b.hi)


def _cell_params_struct(variant: str, ps: Sequence[Tensor]) -> _lib.CellParams:
    names = LSTM_PARAM_NAMES[variant]
    d = {n: _lib.require_cuda_f32(t, n) for n, t in zip(names, ps)}
    p = _lib.CellParams()
    for i, g in enumerate("ifco"):
        p.gcn_kernel[i] = d[f"gcn_kernel_{g}"].data_ptr()
        p.gcn_bias[i] = d[f"gcn_bias_{g}"].data_ptr()
        p.adjacency[i] = (d["adjacency"] if variant == "AGC" else d[f"adjacency_{g}"]).data_ptr()
    if variant != "AAGC":
        for n in ("attention_w", "attention_wq", "attention_wh", "attention_u", "attention_bs", "attention_bu"):
            setattr(p, n, d[n].data_ptr())
    return p


class _LayerTrainFn(torch.autograd.Function):
    """y, hT_0, cT_0[, hT_1, cT_1] = layer(x, (h0, c0) per direction, params per direction)."""

    @staticmethod
    def forward(ctx, meta, x, hmask, *flat):
        variant, nd, reverse, out_act, ws, engine = meta
        names = LSTM_PARAM_NAMES[variant]
        npar = len(names)
        states = flat[:2 * nd]
        params = [flat[2 * nd + d * npar: 2 * nd + (d + 1) * npar] for d in range(nd)]
        x = _lib.require_cuda_f32(x, "input")
        B, T, _, F = x.shape
        H = params[0][0].shape[0]
        dev = x.device
        att = variant != "AAGC"
        f32 = dict(dtype=torch.float32, device=dev)
        tape = {
            "gates": torch.empty(nd, T, B, 4, H, 16, **f32),
            "u": torch.empty(nd, T, B, 4, H, 16, **f32) if variant != "AGC" else None,
            "c": torch.empty(nd, T, B, H, 16, **f32),
            "hh": torch.empty(nd, T, B, H, 16, **f32),
            "e": torch.empty(nd, T, B, H, 16, **f32) if att else None,
            "hp": torch.empty(nd, B, T, NUM_NODES, H, **f32),
            "a": torch.empty(nd, T, B, 16, **f32) if att else None,
            "q": torch.empty(nd, T, B, H, **f32) if att else None,
            "s": torch.empty(nd, T, B, H, **f32) if att else None,
        }
        y = torch.empty(B, T, NUM_NODES, nd * H, **f32)
        h0 = [_lib.require_cuda_f32(states[2 * d], "h0") for d in range(nd)]
        c0 = [_lib.require_cuda_f32(states[2 * d + 1], "c0") for d in range(nd)]
        hT = [torch.empty(B, NUM_NODES, H, **f32) for _ in range(nd)]
        cT = [torch.empty(B, NUM_NODES, H, **f32) for _ in range(nd)]
        cells = (_lib.CellParams * nd)(*[_cell_params_struct(variant, params[d]) for d in range(nd)])
        rev = (C.c_int * nd)(*[int(r) for r in reverse])
        tp = _lib.Tape(*[_lib.ptr(tape[k]) for k in ("gates", "u", "c", "hh", "e", "hp", "a", "q", "s")])
        L = _lib.lib()
        v = _lib.VARIANT[variant]
        with torch.cuda.device(dev):
            wbuf = ws.get(L.a3gc_layer_train_workspace_bytes(v, B, T, F, H, nd, _lib.ENGINE[engine]), dev)
            rc = L.a3gc_layer_train_forward(
                v, nd, cells, rev, x.data_ptr(), T * NUM_NODES * F, NUM_NODES * F,
                _lib.ptr_array(h0, nd), _lib.ptr_array(c0, nd),
                y.data_ptr(), T * NUM_NODES * nd * H, NUM_NODES * nd * H, nd * H,
                _lib.ptr_array(hT, nd), _lib.ptr_array(cT, nd),
                B, T, F, H, _lib.ACT[out_act], C.byref(tp), _lib.ptr(hmask), _lib.ENGINE[engine],
                wbuf.data_ptr(), wbuf.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "a3gc_layer_train_forward")
        ctx.meta, ctx.tape, ctx.shape = meta, tape, (B, T, F, H)
        ctx.save_for_backward(x, hmask if hmask is not None else x.new_empty(0), *h0, *c0, *[t for ps in params for t in ps])
        outs = [y]
        for d in range(nd):
            outs += [hT[d], cT[d]]
        return tuple(outs)

    @staticmethod
    def backward(ctx, dy, *dstates):
        variant, nd, reverse, out_act, ws, engine = ctx.meta
        names = LSTM_PARAM_NAMES[variant]
        npar = len(names)
        B, T, F, H = ctx.shape
        saved = ctx.saved_tensors
        x, hmask = saved[0], (saved[1] if saved[1].numel() else None)
        h0, c0 = saved[2:2 + nd], saved[2 + nd:2 + 2 * nd]
        params = [saved[2 + 2 * nd + d * npar: 2 + 2 * nd + (d + 1) * npar] for d in range(nd)]
        tape = ctx.tape
        if tape is None:
            raise RuntimeError("a3gc_ip_b200: backward through this layer ran twice; the tape is consumed (and overwritten in place) by "
                               "the first pass, retain_graph=True is not supported")
        dev = x.device
        att = variant != "AAGC"
        f32 = dict(dtype=torch.float32, device=dev)
        dy = dy.contiguous() if dy is not None else torch.zeros(B, T, NUM_NODES, nd * H, **f32)
        dhT = [None if dstates[2 * d] is None else dstates[2 * d].contiguous() for d in range(nd)]
        dcT = [None if dstates[2 * d + 1] is None else dstates[2 * d + 1].contiguous() for d in range(nd)]
        gr = {
            "dzm": torch.empty(nd, B, T, NUM_NODES, 4 * H, **f32),
            "dep": torch.empty(nd, T, B, H, 16, **f32) if att else None,
            "dqs": torch.empty(nd, T, B, H, **f32) if att else None,
            "dqp": torch.empty(nd, T, B, H, **f32) if att else None,
            "dap": torch.empty(nd, T, B, 16, **f32) if att else None,
        }
        # mixed mode: the backward writes dzm directly as (TF32-exact head, bf16 head, bf16 remainder) -- no split pass over it
        dzm_mixed = _gemm_mode() == "mixed"
        if dzm_mixed:
            gr["dzm_hi16"] = torch.empty(nd, B, T, NUM_NODES, 4 * H, dtype=torch.bfloat16, device=dev)
            gr["dzm_lo16"] = torch.empty(nd, B, T, NUM_NODES, 4 * H, dtype=torch.bfloat16, device=dev)
        dh0 = [torch.empty(B, NUM_NODES, H, **f32) for _ in range(nd)]
        dc0 = [torch.empty(B, NUM_NODES, H, **f32) for _ in range(nd)]
        cells = (_lib.CellParams * nd)(*[_cell_params_struct(variant, params[d]) for d in range(nd)])
        rev = (C.c_int * nd)(*[int(r) for r in reverse])
        tp = _lib.Tape(*[_lib.ptr(tape[k]) for k in ("gates", "u", "c", "hh", "e", "hp", "a", "q", "s")])
        tg = _lib.TapeGrads(*[_lib.ptr(gr.get(k)) for k in ("dzm", "dep", "dqs", "dqp", "dap", "dzm_hi16", "dzm_lo16")])
        L = _lib.lib()
        v = _lib.VARIANT[variant]
        with torch.cuda.device(dev):
            wbuf = ws.get(L.a3gc_layer_train_workspace_bytes(v, B, T, F, H, nd, _lib.ENGINE[engine]), dev)
            rc = L.a3gc_layer_backward(
                v, nd, cells, rev, dy.data_ptr(), T * NUM_NODES * nd * H, NUM_NODES * nd * H, nd * H,
                _lib.ptr_array(c0, nd), _lib.ptr_array(dhT, nd), _lib.ptr_array(dcT, nd),
                _lib.ptr_array(dh0, nd), _lib.ptr_array(dc0, nd),
                B, T, F, H, _lib.ACT[out_act], C.byref(tp), C.byref(tg), _lib.ptr(hmask),
                wbuf.data_ptr(), wbuf.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "a3gc_layer_backward")

        # ---- hoisted contractions over all (t, b): plain GEMMs (torch / cuBLAS fp32 on the current stream)
        R = B * T * NUM_NODES
        # mixed mode: x and h_prev side by side in one S = [x | h_prev] operand, so dW = dzm^T S is one GEMM per precision
        one_s = _gemm_mode() == "mixed" and F % 4 == 0 and H % 4 == 0
        if one_s:
            S = _mixed_buffers(R, F + H, dev)
            _split_into(S, x.reshape(R, F), 0)
        else:
            x2d = _split(x.reshape(R, F))
        dx = torch.zeros(R, F, **f32)
        grads: List[Optional[Tensor]] = []
        for d in range(nd):
            ps = dict(zip(names, params[d]))
            if dzm_mixed:
                dzm2d = _Split(gr["dzm"][d].reshape(R, 4 * H), None, gr["dzm_hi16"][d].reshape(R, 4 * H), gr["dzm_lo16"][d].reshape(R, 4 * H))
            else:
                dzm2d = _split(gr["dzm"][d].reshape(R, 4 * H))
            # S = [x | h_prev]: h_prev is h' of the previous step of this direction (h0 at its first step), masked
            if one_s:
                _hprev_split(tape["hp"][d], h0[d], None if hmask is None else hmask[d], reverse[d], into=S, col0=F)
                dW = _mm_tn(dzm2d, S)                                    # [4H, F + H]
                gW = [dW[i * H:(i + 1) * H] for i in range(4)]
            else:
                hprev = _hprev_split(tape["hp"][d], h0[d], None if hmask is None else hmask[d], reverse[d])
                dWx = _mm_tn(dzm2d, x2d)                                 # [4H, F]
                dWh = _mm_tn(dzm2d, hprev)                               # [4H, H]
                del hprev
                gW = [torch.cat((dWx[i * H:(i + 1) * H], dWh[i * H:(i + 1) * H]), dim=1) for i in range(4)]
            Wx = torch.cat([ps[f"gcn_kernel_{g}"][:, :F] for g in "ifco"], dim=0)   # [4H, F]
            _addmm_nn(dx, dzm2d, _split(Wx))
            del dzm2d
            dz = tape["gates"][d].reshape(T * B, 4, H, 16)               # the backward left dz here
            ones = torch.ones(1, T * B, **f32)
            gb = (ones @ dz.reshape(T * B, 4 * H * 16)).reshape(4, H, 16).sum(2)     # [4, H]; one bandwidth-bound pass
            out = {f"gcn_kernel_{g}": gW[i] for i, g in enumerate("ifco")}
            out.update({f"gcn_bias_{g}": gb[i] for i, g in enumerate("ifco")})
            if variant == "AGC":
                out["adjacency"] = None                                  # frozen (requires_grad=False, net_aagc.py:238)
            else:
                u = tape["u"][d].reshape(T * B, 4, H, 16)
                # dP_g[m][n] = sum dz_g[., j, m] u_g[., j, n];  adjacency_g is stored as P_g (used as P_g @ S)
                if H % 16 == 0 and os.environ.get("A3GC_TRAIN_ADJ", "fused") == "fused":
                    nblk = 4 * torch.cuda.get_device_properties(dev).multi_processor_count
                    part, dP4 = torch.empty(nblk, 1024, **f32), torch.empty(4, 16, 16, **f32)
                    with torch.cuda.device(dev):
                        rc = L.a3gc_train_adjacency_grad(dz.data_ptr(), u.data_ptr(), T * B, H, part.data_ptr(), nblk, dP4.data_ptr(),
                                                         _lib.stream_ptr(dev))
                    _lib.check(rc, "a3gc_train_adjacency_grad")
                    for i, g in enumerate("ifco"):
                        out[f"adjacency_{g}"] = dP4[i, :NUM_NODES, :NUM_NODES].contiguous()
                else:
                    for i, g in enumerate("ifco"):
                        dP = torch.bmm(dz[:, i].transpose(1, 2), u[:, i]).sum(0)
                        out[f"adjacency_{g}"] = dP[:NUM_NODES, :NUM_NODES].contiguous()
            if att:
                dep = gr["dep"][d].reshape(T * B, H, 16)
                hh = tape["hh"][d].reshape(T * B, H, 16)
                e = tape["e"][d].reshape(T * B, H, 16)
                dap = gr["dap"][d].reshape(T * B, 16)
                # sum over (r, n) of dep[r, k, n] hh[r, j, n]: one [H, R*16] x [R*16, H] GEMM over explicitly permuted copies
                out["attention_wh"] = dep.permute(1, 0, 2).reshape(H, T * B * 16) @ hh.permute(0, 2, 1).reshape(T * B * 16, H)
                out["attention_bs"] = (ones @ dep.reshape(T * B, H * 16)).reshape(H, 16).sum(1)
                out["attention_wq"] = gr["dqs"][d].reshape(T * B, H).t() @ tape["q"][d].reshape(T * B, H)
                out["attention_w"] = gr["dqp"][d].reshape(T * B, H).t() @ tape["s"][d].reshape(T * B, H)
                out["attention_u"] = torch.bmm(e, dap.unsqueeze(2)).sum(0).reshape(1, H)      # sum_{r,n} dap[r,n] e[r,j,n]
                out["attention_bu"] = dap.sum(0)[:NUM_NODES].contiguous()
            grads += [out[n] for n in names]
        state_grads: List[Optional[Tensor]] = []
        for d in range(nd):
            state_grads += [dh0[d], dc0[d]]
        ctx.tape = None
        return (None, dx.reshape(B, T, NUM_NODES, F), None, *state_grads, *grads)


def run_layer_train(variant: str, cells: Sequence[torch.nn.Module], reverse: Sequence[int], x: Tensor,
                    states: Sequence[Tuple[Tensor, Tensor]], out_act: str, ws: _lib.Workspace,
                    p_in: float = 0.0, p_rec: float = 0.0, engine: str = "auto"):
    """Differentiable batch-major forward of one (bi)layer.  x [B,T,15,F] -> (y [B,T,15,nd*H], [(hT, cT)] * nd)."""
    if variant not in LSTM_PARAM_NAMES:
        raise ValueError(f"run_layer_train: unknown LSTM-family variant {variant!r} (the graph-GRU uses run_gru_layer_train)")
    nd = len(cells)
    H = cells[0].units_out
    if x.shape[1] == 0:
        raise RuntimeError("a3gc_ip_b200: the training path needs at least one time step (T == 0)")
    if p_in > 0:
        x = torch.nn.functional.dropout(x, p_in, training=True)                  # net_aagc.py:180
    hmask = None
    if p_rec > 0:
        B, T = x.shape[0], x.shape[1]
        hmask = (torch.rand(nd, B, T, NUM_NODES, H, device=x.device) >= p_rec).to(torch.float32) / (1.0 - p_rec)   # :181
    flat: List[Tensor] = []
    for d in range(nd):
        flat += [states[d][0], states[d][1]]
    for c in cells:
        flat += [getattr(c, n) for n in LSTM_PARAM_NAMES[variant]]
    meta = (variant, nd, tuple(int(r) for r in reverse), out_act, ws, engine)
    outs = _LayerTrainFn.apply(meta, x, hmask, *flat)
    y = outs[0]
    return y, [(outs[1 + 2 * d], outs[2 + 2 * d]) for d in range(nd)]


# ----------------------------------------------------------------------------------------------------------------------
# graph-GRU (G_GRU_cell, net_aagc.py:305-368): same scheme -- tape-keeping forward, reverse-time chain in CUDA, hoisted GEMMs
# ----------------------------------------------------------------------------------------------------------------------
GRU_PARAM_NAMES = ["dense_r_in.weight", "dense_r_in.bias", "dense_u_in.weight", "dense_u_in.bias", "dense_c_in.weight", "dense_c_in.bias",
                   "dense_r_hid.weight", "dense_u_hid.weight", "dense_c_hid.weight", "adjacency", "gcn_kernel"]


def _gru_params_struct(ps: Sequence[Tensor]) -> _lib.CellParams:
    d = {n: _lib.require_cuda_f32(t, n) for n, t in zip(GRU_PARAM_NAMES, ps)}
    p = _lib.CellParams()
    p.g_gcn_kernel, p.g_adjacency = d["gcn_kernel"].data_ptr(), d["adjacency"].data_ptr()
    for i, g in enumerate("ruc"):
        p.dense_in_w[i] = d[f"dense_{g}_in.weight"].data_ptr()
        p.dense_in_b[i] = d[f"dense_{g}_in.bias"].data_ptr()
        p.dense_hid_w[i] = d[f"dense_{g}_hid.weight"].data_ptr()
    return p


class _GruLayerTrainFn(torch.autograd.Function):
    """y, hT_0[, hT_1] = gru_layer(x, h0 per direction, params per direction)."""

    @staticmethod
    def forward(ctx, meta, x, *flat):
        nd, reverse, ws = meta
        npar = len(GRU_PARAM_NAMES)
        h0 = [_lib.require_cuda_f32(t, "h0") for t in flat[:nd]]
        params = [flat[nd + d * npar: nd + (d + 1) * npar] for d in range(nd)]
        x = _lib.require_cuda_f32(x, "input")
        B, T, _, F = x.shape
        H = params[0][-1].shape[0]
        dev = x.device
        f32 = dict(dtype=torch.float32, device=dev)
        tape = {"gates": torch.empty(nd, T, B, 4, H, 16, **f32), "c": torch.empty(nd, T, B, H, 16, **f32),
                "hh": torch.empty(nd, T, B, H, 16, **f32), "hp": torch.empty(nd, B, T, NUM_NODES, H, **f32)}
        y = torch.empty(B, T, NUM_NODES, nd * H, **f32)
        hT = [torch.empty(B, NUM_NODES, H, **f32) for _ in range(nd)]
        cells = (_lib.CellParams * nd)(*[_gru_params_struct(params[d]) for d in range(nd)])
        rev = (C.c_int * nd)(*[int(r) for r in reverse])
        tp = _lib.Tape(_lib.ptr(tape["gates"]), None, _lib.ptr(tape["c"]), _lib.ptr(tape["hh"]), None, _lib.ptr(tape["hp"]), None, None, None)
        L = _lib.lib()
        v = _lib.VARIANT["GGRU"]
        with torch.cuda.device(dev):
            wbuf = ws.get(L.a3gc_layer_train_workspace_bytes(v, B, T, F, H, nd, _lib.ENGINE["simt"]), dev)
            rc = L.a3gc_layer_train_forward(
                v, nd, cells, rev, x.data_ptr(), T * NUM_NODES * F, NUM_NODES * F, _lib.ptr_array(h0, nd), _lib.ptr_array(None, nd),
                y.data_ptr(), T * NUM_NODES * nd * H, NUM_NODES * nd * H, nd * H, _lib.ptr_array(hT, nd), _lib.ptr_array(None, nd),
                B, T, F, H, _lib.ACT["linear"], C.byref(tp), None, _lib.ENGINE["simt"], wbuf.data_ptr(), wbuf.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "a3gc_layer_train_forward")
        ctx.meta, ctx.tape, ctx.shape = meta, tape, (B, T, F, H)
        ctx.save_for_backward(x, *h0, *[t for ps in params for t in ps])
        return (y, *hT)

    @staticmethod
    def backward(ctx, dy, *dhT):
        nd, reverse, ws = ctx.meta
        npar = len(GRU_PARAM_NAMES)
        B, T, F, H = ctx.shape
        saved = ctx.saved_tensors
        x, h0 = saved[0], saved[1:1 + nd]
        params = [saved[1 + nd + d * npar: 1 + nd + (d + 1) * npar] for d in range(nd)]
        tape = ctx.tape
        if tape is None:
            raise RuntimeError("a3gc_ip_b200: backward through this layer ran twice; the tape is consumed (and overwritten in place) by "
                               "the first pass, retain_graph=True is not supported")
        dev = x.device
        f32 = dict(dtype=torch.float32, device=dev)
        dy = dy.contiguous() if dy is not None else torch.zeros(B, T, NUM_NODES, nd * H, **f32)
        dhTc = [None if g is None else g.contiguous() for g in dhT]
        gr = {"dzm": torch.empty(nd, B, T, NUM_NODES, 4 * H, **f32), "dep": torch.empty(nd, T, B, H, 16, **f32),
              "dqs": torch.empty(nd, B, T, NUM_NODES, H, **f32)}
        dh0 = [torch.empty(B, NUM_NODES, H, **f32) for _ in range(nd)]
        cells = (_lib.CellParams * nd)(*[_gru_params_struct(params[d]) for d in range(nd)])
        rev = (C.c_int * nd)(*[int(r) for r in reverse])
        tp = _lib.Tape(_lib.ptr(tape["gates"]), None, _lib.ptr(tape["c"]), _lib.ptr(tape["hh"]), None, _lib.ptr(tape["hp"]), None, None, None)
        tg = _lib.TapeGrads(_lib.ptr(gr["dzm"]), _lib.ptr(gr["dep"]), _lib.ptr(gr["dqs"]), None, None, None, None)
        L = _lib.lib()
        v = _lib.VARIANT["GGRU"]
        with torch.cuda.device(dev):
            wbuf = ws.get(L.a3gc_layer_train_workspace_bytes(v, B, T, F, H, nd, _lib.ENGINE["simt"]), dev)
            rc = L.a3gc_layer_backward(
                v, nd, cells, rev, dy.data_ptr(), T * NUM_NODES * nd * H, NUM_NODES * nd * H, nd * H,
                _lib.ptr_array(list(h0), nd), _lib.ptr_array(dhTc, nd), _lib.ptr_array(None, nd), _lib.ptr_array(dh0, nd), _lib.ptr_array(None, nd),
                B, T, F, H, _lib.ACT["linear"], C.byref(tp), C.byref(tg), None, wbuf.data_ptr(), wbuf.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "a3gc_layer_backward")
        R = B * T * NUM_NODES
        x2d = _split(x.reshape(R, F))
        dx = torch.zeros(R, F, **f32)
        grads: List[Optional[Tensor]] = []
        for d in range(nd):
            ps = dict(zip(GRU_PARAM_NAMES, params[d]))
            db_in = gr["dzm"][d].reshape(R, 4 * H)[:, :3 * H].sum(0)
            dzm2d = _split(gr["dzm"][d].reshape(R, 4 * H))                # (dzr | dzu | dzc | dzc r)
            dz_in = dzm2d.cols(0, 3 * H)
            dW_in = _mm_tn(dz_in, x2d)                                    # [3H, F]
            _addmm_nn(dx, dz_in, _split(torch.cat([ps[f"dense_{g}_in.weight"] for g in "ruc"], dim=0)))
            msg2d = _split(tape["hh"][d].permute(1, 0, 3, 2)[:, :, :NUM_NODES].reshape(R, H))   # [T,B,H,16] -> [B,T,15,H]
            dW_hid = torch.cat((_mm_tn(dzm2d.cols(0, 2 * H), msg2d), _mm_tn(dzm2d.cols(3 * H, 4 * H), msg2d)), dim=0)   # [3H, H]
            del msg2d, dzm2d, dz_in
            hprev = _hprev_split(tape["hp"][d], h0[d], None, reverse[d])
            dWg = _mm_tn(_split(gr["dqs"][d].reshape(R, H)), hprev)       # gcn_kernel [j][k']
            del hprev
            dmsg = gr["dep"][d].reshape(T * B, H, 16)
            M = tape["c"][d].reshape(T * B, H, 16)
            dP = torch.bmm(dmsg.transpose(1, 2), M).sum(0)                # dP[n][m]; the parameter is used transposed (net_aagc.py:348)
            out = {"adjacency": dP.t()[:NUM_NODES, :NUM_NODES].contiguous(), "gcn_kernel": dWg}
            for i, g in enumerate("ruc"):
                out[f"dense_{g}_in.weight"] = dW_in[i * H:(i + 1) * H]
                out[f"dense_{g}_in.bias"] = db_in[i * H:(i + 1) * H]
                out[f"dense_{g}_hid.weight"] = dW_hid[i * H:(i + 1) * H]
            grads += [out[n] for n in GRU_PARAM_NAMES]
        ctx.tape = None
        return (None, dx.reshape(B, T, NUM_NODES, F), *dh0, *grads)


def run_gru_layer_train(cells: Sequence[torch.nn.Module], reverse: Sequence[int], x: Tensor, states: Sequence[Tensor], ws: _lib.Workspace):
    """Differentiable batch-major forward of one (bi) graph-GRU layer.  The cell's dropout arguments are unused (net_aagc.py:343-368)."""
    nd = len(cells)
    if x.shape[1] == 0:
        raise RuntimeError("a3gc_ip_b200: the training path needs at least one time step (T == 0)")
    flat: List[Tensor] = list(states)
    for c in cells:
        for n in GRU_PARAM_NAMES:
            obj = c
            for part in n.split("."):
                obj = getattr(obj, part)
            flat.append(obj)
    outs = _GruLayerTrainFn.apply((nd, tuple(int(r) for r in reverse), ws), x, *flat)
    return outs[0], list(outs[1:])


def gc_train(mod: torch.nn.Module, x: Tensor, act: str, p_drop: float) -> Tensor:
    """AAGC.forward (net_aagc.py:61-66) with autograd history: the non-recurrent graph convolution is <1 % of the
    step's FLOPs; in the training path it is expressed with torch ops (einsum + matmul) so autograd provides its
    backward."""
    if p_drop > 0:
        x = torch.nn.functional.dropout(x, p_drop, training=True)
    # (adj @ x) @ W^T = adj @ (x @ W^T): contract the wide side first, so the node mix touches min(F, O) features
    if mod.gcn_kernel.shape[0] < mod.gcn_kernel.shape[1]:
        y = torch.einsum("bsnf,nm->bsmf", torch.matmul(x, mod.gcn_kernel.t()), mod.adj.t()).contiguous() + mod.gcn_bias
    else:
        y = torch.matmul(torch.einsum("bsnf,nm->bsmf", x, mod.adj.t()), mod.gcn_kernel.t()) + mod.gcn_bias
    if act == "tanh":
        y = torch.tanh(y)
    elif act == "relu":
        y = torch.relu(y)
    return y
