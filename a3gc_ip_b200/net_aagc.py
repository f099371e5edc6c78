"""Drop-in replacements for the recurrent graph-convolution classes of the reference's
``net_aagc.py`` (lines 40-695): same class names, constructor arguments, ``forward`` signatures,
parameter names / shapes / registration order (so the reference's ``state_dict``s -- e.g.
``trained_models/A3GC/*.tar`` -- load with ``strict=True``), and the same ``ValueError`` /
``AssertionError`` behaviour.  ``forward`` runs hand-written sm_100a CUDA through the C ABI in
``include/a3gc_b200.h``; there is no CPU or eager-PyTorch fallback.

Differences from the reference, all deliberate:
  * adjacency parameters are cloned, never aliased to the caller's template (the reference's
    ``Parameter(adjacency_matrix.t())`` shares one buffer between all cells on CPU; SURVEY #3);
  * the forward and reverse directions of a ``Bi*`` layer run concurrently in one launch;
  * ``.eval()`` mode is the inference path (tensor-core engine, no autograd history); ``.train()`` mode is the
    training path of the LSTM-family classes (a3gc_ip_b200/training.py): forward keeps a tape, backward runs
    the BPTT chain in CUDA, dropout masks are drawn per call (the graph-GRU cell has no dropout in the reference either).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor
from torch.nn import Parameter

from . import _lib
from . import training as _tr

NUM_NODES = 15


def _check_activation(activation_fn: str) -> str:
    if activation_fn not in ("linear", "tanh"):
        raise ValueError("only support linear and tanh activations for now")     # net_aagc.py:51
    return activation_fn


def _adj_param(adjacency_matrix: Tensor, requires_grad: bool = True, transpose: bool = True) -> Parameter:
    if tuple(adjacency_matrix.shape) != (NUM_NODES, NUM_NODES):
        raise RuntimeError(f"the B200 kernels support the 15-node skeleton graph only, got adjacency {tuple(adjacency_matrix.shape)}")
    a = adjacency_matrix.t() if transpose else adjacency_matrix
    return Parameter(a.detach().to(torch.float32).clone().contiguous(), requires_grad=requires_grad)


class _EngineMixin:
    """Engine / precision selection shared by all modules (not part of the reference's API)."""
    engine: str = os.environ.get("A3GC_ENGINE", "auto")
    precision: str = os.environ.get("A3GC_PRECISION", "fp32")

    def set_engine(self, engine: str = "auto", precision: str = "fp32"):
        if engine not in _lib.ENGINE or precision not in _lib.PRECISION:
            raise ValueError(f"engine in {list(_lib.ENGINE)}, precision in {list(_lib.PRECISION)}")
        for m in self.modules():
            if isinstance(m, _EngineMixin):
                m.engine, m.precision = engine, precision
        return self

    def release_workspaces(self) -> None:
        """Drop the scratch buffers (operand images, inter-layer activations; tens of GB at BASELINE cfg 2) held by this
        module and its children; ``.cpu()`` / ``.to()`` do not move or free them, the next forward re-allocates."""
        for m in self.modules():
            for v in list(vars(m).values()):
                if isinstance(v, _lib.Workspace):
                    v.release()
                elif isinstance(v, dict):
                    for w in v.values():
                        if isinstance(w, _lib.Workspace):
                            w.release()


# ----------------------------------------------------------------------------------------
# AAGC graph convolution (net_aagc.py:40-66)
# ----------------------------------------------------------------------------------------
class AAGC(torch.nn.Module, _EngineMixin):
    """First X*W, then A*X. A learnable.  (net_aagc.py:40-66)"""

    def __init__(self, units_in, units_out, adjacency_matrix, activation_fn="linear", dropout=0.0):
        super().__init__()
        self.activation_name = _check_activation(activation_fn)
        self.p_dropout = float(dropout)
        self.dropout = torch.nn.Dropout(dropout)
        self.gcn_kernel = Parameter(torch.zeros((units_out, units_in), dtype=torch.float32))
        self.adj = _adj_param(adjacency_matrix)
        self.gcn_bias = Parameter(torch.zeros(units_out, dtype=torch.float32))
        torch.nn.init.xavier_uniform_(self.gcn_kernel)

    def _params(self) -> _lib.GcParams:
        for p in (self.gcn_kernel, self.adj, self.gcn_bias):
            _lib.require_cuda_f32(p.data, "AAGC parameter")
        return _lib.GcParams(self.gcn_kernel.data_ptr(), self.adj.data_ptr(), self.gcn_bias.data_ptr())

    def forward(self, input: Tensor, _act: Optional[str] = None) -> Tensor:
        if self.training:
            return _tr.gc_train(self, _lib.require_cuda_f32(input, "input"), _act or self.activation_name, self.p_dropout)
        x = _lib.require_cuda_f32(input, "input")
        if x.dim() != 4 or x.shape[2] != NUM_NODES or x.shape[3] != self.gcn_kernel.shape[1]:
            raise RuntimeError(f"AAGC expects [B, T, 15, {self.gcn_kernel.shape[1]}], got {tuple(x.shape)}")
        B, T = x.shape[0], x.shape[1]
        f_out = self.gcn_kernel.shape[0]
        y = torch.empty(B, T, NUM_NODES, f_out, dtype=torch.float32, device=x.device)
        p = self._params()
        with torch.cuda.device(x.device):
            rc = _lib.lib().a3gc_gc_forward(C.byref(p), x.data_ptr(), y.data_ptr(), B * T, x.shape[3], f_out,
                                            _lib.ACT[_act or self.activation_name], _lib.stream_ptr(x.device))
        _lib.check(rc, "a3gc_gc_forward")
        return y


# ----------------------------------------------------------------------------------------
# cells (net_aagc.py:68-368)
# ----------------------------------------------------------------------------------------
class _CellBase(torch.nn.Module, _EngineMixin):
    variant: str = ""

    def _cell_params(self) -> _lib.CellParams:
        raise NotImplementedError

    @property
    def units_out(self) -> int:
        raise NotImplementedError

    @property
    def units_in(self) -> int:
        raise NotImplementedError


def _run_layer(variant: str, cells: Sequence[_CellBase], reverse: Sequence[int], x: Tensor, time_major: bool,
               states: Sequence, out_act: str, ws: _lib.Workspace, engine: str, precision: str):
    """Common host path of cell / layer / bi-layer forwards -> a3gc_layer_forward.

    x: [T,B,15,F] if time_major else [B,T,15,F].  states[d]: (h, c) for the LSTM family, h for G-GRU.
    Returns (y, out_states) with y [.., .., 15, len(cells)*H] in the same major order as x.
    """
    x = _lib.require_cuda_f32(x, "input")
    nd = len(cells)
    H, F = cells[0].units_out, cells[0].units_in
    if x.dim() != 4 or x.shape[2] != NUM_NODES or x.shape[3] != F:
        raise RuntimeError(f"{type(cells[0]).__name__} expects [*, *, 15, {F}] input, got {tuple(x.shape)}")
    if time_major:
        T, B = x.shape[0], x.shape[1]
        sxb, sxt = NUM_NODES * F, B * NUM_NODES * F
        y = torch.empty(T, B, NUM_NODES, nd * H, dtype=torch.float32, device=x.device)
        syb, syt = NUM_NODES * nd * H, B * NUM_NODES * nd * H
    else:
        B, T = x.shape[0], x.shape[1]
        sxb, sxt = T * NUM_NODES * F, NUM_NODES * F
        y = torch.empty(B, T, NUM_NODES, nd * H, dtype=torch.float32, device=x.device)
        syb, syt = T * NUM_NODES * nd * H, NUM_NODES * nd * H
    gru = variant == "GGRU"
    h0, c0 = [], []
    for d in range(nd):
        s = states[d]
        h, c = (s, None) if gru else s
        for t in ((h,) if gru else (h, c)):
            if tuple(t.shape) != (B, NUM_NODES, H):
                raise RuntimeError(f"state must be [{B}, 15, {H}], got {tuple(t.shape)}")
        h0.append(_lib.require_cuda_f32(h, "state"))
        c0.append(None if gru else _lib.require_cuda_f32(c, "state"))
    hT = [torch.empty(B, NUM_NODES, H, dtype=torch.float32, device=x.device) for _ in range(nd)]
    cT = [None if gru else torch.empty(B, NUM_NODES, H, dtype=torch.float32, device=x.device) for _ in range(nd)]
    params = (_lib.CellParams * nd)(*[c._cell_params() for c in cells])
    rev = (C.c_int * nd)(*[int(r) for r in reverse])
    L = _lib.lib()
    v, pr, en = _lib.VARIANT[variant], _lib.PRECISION[precision], _lib.ENGINE[engine]
    with torch.cuda.device(x.device):
        nbytes = L.a3gc_layer_workspace_bytes(v, B, T, F, H, nd, pr, en)
        wbuf = ws.get(nbytes, x.device)
        rc = L.a3gc_layer_forward(v, nd, params, rev, x.data_ptr(), sxb, sxt,
                                  _lib.ptr_array(h0, nd), _lib.ptr_array(c0, nd),
                                  y.data_ptr(), syb, syt, nd * H,
                                  _lib.ptr_array(hT, nd), _lib.ptr_array(cT, nd),
                                  B, T, F, H, _lib.ACT[out_act], pr, en,
                                  wbuf.data_ptr(), wbuf.numel(), _lib.stream_ptr(x.device))
    _lib.check(rc, "a3gc_layer_forward")
    out_states = [hT[d] if gru else (hT[d], cT[d]) for d in range(nd)]
    return y, out_states


class _LSTMCellBase(_CellBase):
    has_attention = False
    single_adjacency = False

    def __init__(self, units_in, units_out, adjacency_matrix, activation_fn="linear", dropout=0.0, recurrent_dropout=0.0):
        super().__init__()
        self.activation_name = _check_activation(activation_fn)
        if self.has_attention or self.single_adjacency:
            num_nodes = adjacency_matrix.shape[-1]
            assert num_nodes == 15                                                # net_aagc.py:141-142
        self.p_dropout, self.p_recurrent_dropout = float(dropout), float(recurrent_dropout)
        self.dropout = torch.nn.Dropout(dropout)
        self.recurrent_dropout = torch.nn.Dropout(recurrent_dropout)
        mk = lambda: Parameter(torch.zeros((units_out, units_in + units_out), dtype=torch.float32))
        if self.single_adjacency:                                                 # AGC registers adjacency first (:238)
            self.adjacency = _adj_param(adjacency_matrix, requires_grad=False)
        for g in "ifco":
            setattr(self, f"gcn_kernel_{g}", mk())
        if not self.single_adjacency:
            for g in "ifco":
                setattr(self, f"adjacency_{g}", _adj_param(adjacency_matrix))
        for g in "ifco":
            setattr(self, f"gcn_bias_{g}", Parameter(torch.zeros(units_out, dtype=torch.float32)))
        if self.has_attention:
            self.attention_w = Parameter(torch.zeros((units_out, units_out), dtype=torch.float32))
            self.attention_wq = Parameter(torch.zeros((units_out, units_out), dtype=torch.float32))
            self.attention_wh = Parameter(torch.zeros((units_out, units_out), dtype=torch.float32))
            self.attention_u = Parameter(torch.zeros((1, units_out), dtype=torch.float32))
            self.attention_bs = Parameter(torch.zeros(units_out, dtype=torch.float32))
            self.attention_bu = Parameter(torch.zeros(NUM_NODES, dtype=torch.float32))
        for g in "ifco":
            torch.nn.init.xavier_uniform_(getattr(self, f"gcn_kernel_{g}"))
        if self.has_attention:
            for n in ("attention_w", "attention_wq", "attention_wh", "attention_u"):
                torch.nn.init.xavier_uniform_(getattr(self, n))
        self._ws = _lib.Workspace()

    @property
    def units_out(self) -> int:
        return self.gcn_kernel_i.shape[0]

    @property
    def units_in(self) -> int:
        return self.gcn_kernel_i.shape[1] - self.gcn_kernel_i.shape[0]

    def _cell_params(self) -> _lib.CellParams:
        p = _lib.CellParams()
        for i, g in enumerate("ifco"):
            p.gcn_kernel[i] = _lib.require_cuda_f32(getattr(self, f"gcn_kernel_{g}").data, "parameter").data_ptr()
            p.gcn_bias[i] = _lib.require_cuda_f32(getattr(self, f"gcn_bias_{g}").data, "parameter").data_ptr()
            adj = self.adjacency if self.single_adjacency else getattr(self, f"adjacency_{g}")
            p.adjacency[i] = _lib.require_cuda_f32(adj.data, "parameter").data_ptr()
        if self.has_attention:
            for n in ("attention_w", "attention_wq", "attention_wh", "attention_u", "attention_bs", "attention_bu"):
                setattr(p, n, _lib.require_cuda_f32(getattr(self, n).data, "parameter").data_ptr())
        return p

    def forward(self, input: Tensor, state: Tuple[Tensor, Tensor]) -> Tuple[Tensor, Tuple[Tensor, Tensor]]:
        """One step: input [B,15,F], state (h, c) [B,15,H] -> (act(h'), (h', c'))."""
        if self.training:
            y, st = _tr.run_layer_train(self.variant, [self], [0], input.unsqueeze(1), [state], self.activation_name, self._ws,
                                        self.p_dropout, self.p_recurrent_dropout, self.engine)
            return y[:, 0], st[0]
        y, st = _run_layer(self.variant, [self], [0], input.unsqueeze(0), True, [state], self.activation_name,
                           self._ws, self.engine, self.precision)
        return y[0], st[0]


class AAGC_LSTM_cell(_LSTMCellBase):
    """net_aagc.py:68-126"""
    variant = "AAGC"


class A3GC_LSTM_cell(_LSTMCellBase):
    """net_aagc.py:128-217"""
    variant = "A3GC"
    has_attention = True


class AGC_LSTM_cell(_LSTMCellBase):
    """net_aagc.py:219-303"""
    variant = "AGC"
    has_attention = True
    single_adjacency = True


class G_GRU_cell(_CellBase):
    """net_aagc.py:305-368.  ``activation_fn`` / dropout arguments are accepted and unused, as in the reference."""
    variant = "GGRU"

    def __init__(self, units_in, units_out, adjacency_matrix, activation_fn="linear", dropout=0.0, recurrent_dropout=0.0):
        super().__init__()
        self.activation_name = _check_activation(activation_fn)
        num_nodes = adjacency_matrix.shape[-1]
        assert num_nodes == 15                                                    # :318-319
        self.a = _adj_param(adjacency_matrix, requires_grad=False, transpose=False)   # frozen, unused (:324)
        self.dense_r_in = torch.nn.Linear(units_in, units_out, bias=True)
        self.dense_u_in = torch.nn.Linear(units_in, units_out, bias=True)
        self.dense_c_in = torch.nn.Linear(units_in, units_out, bias=True)
        self.dense_r_hid = torch.nn.Linear(units_out, units_out, bias=False)
        self.dense_u_hid = torch.nn.Linear(units_out, units_out, bias=False)
        self.dense_c_hid = torch.nn.Linear(units_out, units_out, bias=False)
        self.adjacency = _adj_param(adjacency_matrix)
        self.gcn_kernel = Parameter(torch.zeros((units_out, units_out), dtype=torch.float32))
        torch.nn.init.xavier_uniform_(self.adjacency)                             # :339
        torch.nn.init.xavier_uniform_(self.gcn_kernel)
        self._ws = _lib.Workspace()

    @property
    def units_out(self) -> int:
        return self.gcn_kernel.shape[0]

    @property
    def units_in(self) -> int:
        return self.dense_r_in.weight.shape[1]

    def _cell_params(self) -> _lib.CellParams:
        p = _lib.CellParams()
        q = lambda t: _lib.require_cuda_f32(t.data, "parameter").data_ptr()
        p.g_gcn_kernel, p.g_adjacency = q(self.gcn_kernel), q(self.adjacency)
        for i, g in enumerate("ruc"):
            p.dense_in_w[i] = q(getattr(self, f"dense_{g}_in").weight)
            p.dense_in_b[i] = q(getattr(self, f"dense_{g}_in").bias)
            p.dense_hid_w[i] = q(getattr(self, f"dense_{g}_hid").weight)
        return p

    def forward(self, input: Tensor, state: Tensor) -> Tuple[Tensor, Tensor]:
        if self.training:
            y, st = _tr.run_gru_layer_train([self], [0], input.unsqueeze(1), [state], self._ws)
            return st[0], st[0]
        y, st = _run_layer("GGRU", [self], [0], input.unsqueeze(0), True, [state], "linear", self._ws, self.engine, self.precision)
        return st[0], st[0]


# ----------------------------------------------------------------------------------------
# layers (net_aagc.py:370-592)
# ----------------------------------------------------------------------------------------
class _Layer(torch.nn.Module, _EngineMixin):
    """Time loop over a time-major input [T, B, 15, F]; one direction."""
    cell_cls = None
    reverse = 0

    def __init__(self, *cell_args, **cell_kwargs):
        super().__init__()
        self.cell = self.cell_cls(*cell_args, **cell_kwargs)
        self._ws = _lib.Workspace()

    def forward(self, input: Tensor, state):
        c = self.cell
        act = "linear" if c.variant == "GGRU" else c.activation_name
        if self.training:
            if c.variant == "GGRU":
                y, st = _tr.run_gru_layer_train([c], [self.reverse], input.transpose(0, 1).contiguous(), [state], self._ws)
                return y.transpose(0, 1), st[0]
            y, st = _tr.run_layer_train(c.variant, [c], [self.reverse], input.transpose(0, 1).contiguous(), [state], act, self._ws,
                                        c.p_dropout, c.p_recurrent_dropout, self.engine)
            return y.transpose(0, 1), st[0]
        y, st = _run_layer(c.variant, [c], [self.reverse], input, True, [state], act, self._ws, self.engine, self.precision)
        return y, st[0]


class _BiLayer(torch.nn.Module, _EngineMixin):
    """Both directions over a batch-major input [B, T, 15, F] -> [B, T, 15, 2H]  (net_aagc.py:469-480)."""
    fwd_cls = None
    rev_cls = None

    def __init__(self, *cell_args, **cell_kwargs):
        super().__init__()
        self.directions = torch.nn.ModuleList([self.fwd_cls(*cell_args, **cell_kwargs), self.rev_cls(*cell_args, **cell_kwargs)])
        self._ws = _lib.Workspace()

    def forward(self, input: Tensor, states: List):
        cells = [d.cell for d in self.directions]
        c = cells[0]
        act = "linear" if c.variant == "GGRU" else c.activation_name
        if self.training:
            if c.variant == "GGRU":
                return _tr.run_gru_layer_train(cells, [0, 1], input, states, self._ws)
            return _tr.run_layer_train(c.variant, cells, [0, 1], input, states, act, self._ws, c.p_dropout, c.p_recurrent_dropout,
                                       self.engine)
        return _run_layer(c.variant, cells, [0, 1], input, False, states, act, self._ws, self.engine, self.precision)


class AAGC_LSTM(_Layer):
    cell_cls = AAGC_LSTM_cell


class ReverseAAGC_LSTM(_Layer):
    cell_cls = AAGC_LSTM_cell
    reverse = 1


class BiAAGC_LSTM(_BiLayer):
    fwd_cls, rev_cls = AAGC_LSTM, ReverseAAGC_LSTM


class A3GC_LSTM(_Layer):
    cell_cls = A3GC_LSTM_cell


class ReverseA3GC_LSTM(_Layer):
    cell_cls = A3GC_LSTM_cell
    reverse = 1


class BiA3GC_LSTM(_BiLayer):
    fwd_cls, rev_cls = A3GC_LSTM, ReverseA3GC_LSTM


class AGC_LSTM(_Layer):
    cell_cls = AGC_LSTM_cell


class ReverseAGC_LSTM(_Layer):
    cell_cls = AGC_LSTM_cell
    reverse = 1


class BiAGC_LSTM(_BiLayer):
    fwd_cls, rev_cls = AGC_LSTM, ReverseAGC_LSTM


class G_GRU(_Layer):
    cell_cls = G_GRU_cell


class ReverseG_GRU(_Layer):
    cell_cls = G_GRU_cell
    reverse = 1


class BiG_GRU(_BiLayer):
    fwd_cls, rev_cls = G_GRU, ReverseG_GRU


# ----------------------------------------------------------------------------------------
# nets (net_aagc.py:595-695)
# ----------------------------------------------------------------------------------------
class _Net(torch.nn.Module, _EngineMixin):
    r"""linear_in -> relu -> rnn1 -> rnn2 (seeded with rnn1's final state) -> linear_out."""
    variant = ""
    bi_cls = None

    def __init__(self, units_in, units_out, units_hidden, adjacency_matrix, linear_dropout=0.2, dropout=0.3, recurrent_dropout=0.3):
        super().__init__()
        self.units_hidden = units_hidden
        self.linear_in = AAGC(units_in, units_hidden, adjacency_matrix, activation_fn="linear", dropout=linear_dropout)
        # the reference passes recurrent_dropout=dropout and ignores the net's own argument (net_aagc.py:629-630)
        self.rnn1 = self.bi_cls(units_hidden, units_hidden, adjacency_matrix, activation_fn="tanh", dropout=dropout, recurrent_dropout=dropout)
        self.rnn2 = self.bi_cls(units_hidden * 2, units_hidden, adjacency_matrix, activation_fn="tanh", dropout=dropout, recurrent_dropout=dropout)
        self.linear_out = AAGC(units_hidden * 2, units_out, adjacency_matrix, activation_fn="linear", dropout=0.0)
        self._ws = _lib.Workspace()
        self._pack_cache = False

    def cache_packed_weights(self, on: bool = True):
        """Opt in to keeping the packed (operand-image) weights of rnn1 / rnn2 across calls instead of re-packing them in
        every forward (~25 us per layer, i.e. about 1 % of a B = 1 call).  For serving with frozen weights.  The
        cache is keyed on every parameter's storage pointer and autograd version, so ``load_state_dict``, optimizer steps
        and ``p.data = ...`` invalidate it; in-place edits through ``p.data`` that keep the storage do not bump the version
        -- call ``invalidate_packed_weights()`` after those (the default, no cache, is always exact)."""
        self._pack_cache = bool(on)
        self.invalidate_packed_weights()
        return self

    def invalidate_packed_weights(self) -> None:
        self.__dict__.pop("_packed", None)

    def _packed_ptrs(self, dev):
        """[ptr or None] * 2 for rnn1, rnn2: packs on the current stream when the key changed, and makes the current
        stream wait for the packing stream otherwise."""
        if not self._pack_cache:
            return [None, None]
        L = _lib.lib()
        v, pr, en = _lib.VARIANT[self.variant], _lib.PRECISION[self.precision], _lib.ENGINE[self.engine]
        H = self.units_hidden
        cache = self.__dict__.setdefault("_packed", {})
        out = []
        for l, (rnn, f_in) in enumerate(((self.rnn1, H), (self.rnn2, 2 * H))):
            key = (pr, en, str(dev)) + tuple((q.data_ptr(), q._version) for q in rnn.parameters())
            ent = cache.get(l)
            if ent is None or ent[0] != key:
                with torch.cuda.device(dev):
                    nbytes = L.a3gc_packed_weights_bytes(v, f_in, H, 2, pr, en)
                    if nbytes == 0:
                        ent = (key, None, None)
                    else:
                        buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                        cells = (_lib.CellParams * 2)(*[rnn.directions[d].cell._cell_params() for d in range(2)])
                        rc = L.a3gc_pack_weights(v, 2, cells, f_in, H, pr, en, buf.data_ptr(), nbytes, _lib.stream_ptr(dev))
                        _lib.check(rc, "a3gc_pack_weights")
                        ev = torch.cuda.Event()
                        ev.record(torch.cuda.current_stream(dev))
                        ent = (key, buf, ev)
                cache[l] = ent
            elif ent[2] is not None:
                torch.cuda.current_stream(dev).wait_event(ent[2])     # packed on another stream (concurrent batch chunks)
            out.append(None if ent[1] is None else ent[1].data_ptr())
        return out

    def _workspace(self, slot: int) -> "_lib.Workspace":
        if slot == 0:
            return self._ws
        pool = self.__dict__.setdefault("_ws_pool", {})
        if slot not in pool:
            pool[slot] = _lib.Workspace()
        return pool[slot]

    def _net_params(self) -> _lib.NetParams:
        p = _lib.NetParams()
        p.linear_in = self.linear_in._params()
        p.linear_out = self.linear_out._params()
        for l, rnn in enumerate((self.rnn1, self.rnn2)):
            for d in range(2):
                p.rnn[l][d] = rnn.directions[d].cell._cell_params()
        return p

    def forward(self, x: Tensor, h=None, _slot: int = 0, out: Optional[Tensor] = None):
        """x [B, T, 15, units_in] -> (y [B, T, 15, units_out], rnn2 final states).

        ``_slot`` (not part of the reference's signature) selects one of the module's workspaces, so that calls on
        different CUDA streams (batch chunks run concurrently by ``TPPipeline``) do not share scratch memory; ``out``
        (optional) receives y.

        h: None (zero state) or the reference's structure: ``[(h, c), (h, c)]`` ([h, h] for G-GRU),
        each [B, 15, units_hidden]; the returned states have the same structure.
        """
        if self.training:
            return self._forward_train(x, h)
        x = _lib.require_cuda_f32(x, "x")
        f0 = self.linear_in.gcn_kernel.shape[1]
        if x.dim() != 4 or x.shape[2] != NUM_NODES or x.shape[3] != f0:
            raise RuntimeError(f"{type(self).__name__} expects x of shape [B, T, 15, {f0}], got {tuple(x.shape)}")
        return self._run(x.shape[0], x.shape[1], x.device, h, _slot, out, x=x)

    @torch.no_grad()
    def forward_raw(self, acc: Tensor, ori: Tensor, stats: Optional[dict] = None, pos: Optional[Tensor] = None, h=None,
                    _slot: int = 0, out: Optional[Tensor] = None):
        """The net fed with the raw IMU frame: ``prepare_input`` (evaluate_a3gc_tp.py:64-94) and, with ``pos``, the
        stage concatenation ``torch.cat((x, pos.view(B, T, 15, 3)), dim=-1)`` (:168, :170) are fused into the load of
        ``linear_in`` (``a3gc_net_forward_raw``).  acc [B, T, 18], ori [B, T, 54]; ``stats`` = the dict of
        ``data/all*_train_stats.pt`` (``--norm``) or None; units_in must be 12 (no pos) or 15 (pos [B, T, 15, 3])."""
        if self.training:
            raise RuntimeError("forward_raw is the inference path (call .eval() first)")
        acc = _lib.require_cuda_f32(acc, "acc")
        ori = _lib.require_cuda_f32(ori, "ori")
        if acc.dim() != 3 or ori.dim() != 3 or acc.shape[-1] != 18 or ori.shape[-1] != 54 or acc.shape[:2] != ori.shape[:2]:
            raise RuntimeError(f"forward_raw expects acc [B,T,18] and ori [B,T,54], got {tuple(acc.shape)} / {tuple(ori.shape)}")
        B, T = acc.shape[0], acc.shape[1]
        f0 = self.linear_in.gcn_kernel.shape[1]
        if f0 != (12 if pos is None else 15):
            raise RuntimeError(f"forward_raw: units_in = {f0} does not match the raw input ({'12 without' if pos is None else '15 with'} pos)")
        if pos is not None:
            pos = _lib.require_cuda_f32(pos, "pos")
            if pos.numel() != B * T * NUM_NODES * 3:
                raise RuntimeError(f"pos must hold [B, T, 15, 3], got {tuple(pos.shape)}")
        return self._run(B, T, acc.device, h, _slot, out, raw=(acc, ori, _stats_on(stats, acc.device), pos))

    def _run(self, B, T, dev, h, _slot, out, x=None, raw=None):
        gru = self.variant == "GGRU"
        f0, H, O = self.linear_in.gcn_kernel.shape[1], self.units_hidden, self.linear_out.gcn_kernel.shape[0]
        h0 = c0 = None
        if h is not None:
            if len(h) != 2:
                raise RuntimeError("h must hold one state per direction")
            hs = [s if gru else s[0] for s in h]
            for t in hs:
                if tuple(t.shape) != (B, NUM_NODES, H):
                    raise RuntimeError(f"state must be [{B}, 15, {H}], got {tuple(t.shape)}")
            h0 = [_lib.require_cuda_f32(t, "h") for t in hs]
            c0 = None if gru else [_lib.require_cuda_f32(s[1], "c") for s in h]
        if out is not None:
            if not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (B, T, NUM_NODES, O)):
                raise RuntimeError(f"out must be a contiguous CUDA float32 tensor of shape [{B}, {T}, 15, {O}]")
            y = out
        else:
            y = torch.empty(B, T, NUM_NODES, O, dtype=torch.float32, device=dev)
        hT = [torch.empty(B, NUM_NODES, H, dtype=torch.float32, device=dev) for _ in range(2)]
        cT = None if gru else [torch.empty(B, NUM_NODES, H, dtype=torch.float32, device=dev) for _ in range(2)]
        p = self._net_params()
        packed = self._packed_ptrs(dev)
        p.packed_rnn[0], p.packed_rnn[1] = packed[0], packed[1]
        L = _lib.lib()
        v, pr, en = _lib.VARIANT[self.variant], _lib.PRECISION[self.precision], _lib.ENGINE[self.engine]
        with torch.cuda.device(dev):
            nbytes = L.a3gc_net_workspace_bytes(v, B, T, f0, H, O, pr, en)
            if nbytes == 0 and B * T > 0:
                raise RuntimeError("a3gc_net_workspace_bytes: " + L.a3gc_last_error().decode(errors="replace"))
            wbuf = self._workspace(_slot).get(nbytes, dev)
            if raw is None:
                rc = L.a3gc_net_forward(v, C.byref(p), x.data_ptr(), _lib.ptr_array(h0), _lib.ptr_array(c0), y.data_ptr(),
                                        _lib.ptr_array(hT), _lib.ptr_array(cT), B, T, f0, H, O, pr, en,
                                        wbuf.data_ptr(), wbuf.numel(), _lib.stream_ptr(dev))
            else:
                acc, ori, st, pos = raw
                rc = L.a3gc_net_forward_raw(v, C.byref(p), acc.data_ptr(), ori.data_ptr(), *(st or (None,) * 4), _lib.ptr(pos),
                                            _lib.ptr_array(h0), _lib.ptr_array(c0), y.data_ptr(),
                                            _lib.ptr_array(hT), _lib.ptr_array(cT), B, T, H, O, pr, en,
                                            wbuf.data_ptr(), wbuf.numel(), _lib.stream_ptr(dev))
                self._keep = st                      # the statistics vectors must outlive the enqueued kernels
        _lib.check(rc, "a3gc_net_forward" if raw is None else "a3gc_net_forward_raw")
        h_out = [hT[0], hT[1]] if gru else [(hT[0], cT[0]), (hT[1], cT[1])]
        return y, h_out


_STATS_CACHE = {}


def _stats_on(stats: Optional[dict], dev):
    """(acc_mean, acc_std, ori_mean, ori_std) pointers of a ``data/all*_train_stats.pt`` dict on ``dev`` (cached per dict)."""
    if stats is None:
        return None
    key = (id(stats), str(dev))
    hit = _STATS_CACHE.get(key)
    if hit is None:
        ts = tuple(stats[a][b].to(dev, torch.float32).contiguous() for a, b in
                   (("acc", "mean_channel"), ("acc", "std_channel"), ("ori", "mean_channel"), ("ori", "std_channel")))
        hit = (stats, ts)                             # keeps `stats` alive so its id() cannot be recycled
        _STATS_CACHE[key] = hit
    return _StatPtrs(hit[1])


class _StatPtrs(tuple):
    """Four device tensors that unpack to their data pointers in a ctypes call."""

    def __new__(cls, ts):
        self = super().__new__(cls, [t.data_ptr() for t in ts])
        self.tensors = ts
        return self


def _net_forward_train(self, x: Tensor, h=None):
    """Training-mode forward of the nets (net_aagc.py:633-645 under ``model.train()``, train_a3gc_tp.py:74): same
    chain as the inference path, each stage differentiable; dropout as the reference configures it."""
    x = _lib.require_cuda_f32(x, "x")
    B, H = x.shape[0], self.units_hidden
    if h is None:
        z = lambda: torch.zeros(B, NUM_NODES, H, dtype=torch.float32, device=x.device)     # net_aagc.py:634-639, :686-687
        h = [z(), z()] if self.variant == "GGRU" else [(z(), z()), (z(), z())]
    a = torch.relu(self.linear_in(x))
    a, h = self.rnn1(a, h)
    a, h = self.rnn2(a, h)
    return self.linear_out(a), h


_Net._forward_train = _net_forward_train


class AAGC_net(_Net):
    """net_aagc.py:595-619"""
    variant, bi_cls = "AAGC", BiAAGC_LSTM


class A3GC_net(_Net):
    """net_aagc.py:621-645"""
    variant, bi_cls = "A3GC", BiA3GC_LSTM


class AGC_net(_Net):
    """net_aagc.py:647-671"""
    variant, bi_cls = "AGC", BiAGC_LSTM


class G_GRU_net(_Net):
    """net_aagc.py:673-695"""
    variant, bi_cls = "GGRU", BiG_GRU


# ----------------------------------------------------------------------------------------
# pipeline wrappers (net_aagc.py:697-965): SMPL-free shims so 'pose_net.*' checkpoints load strict
# ----------------------------------------------------------------------------------------
class _PoseNetBase(torch.nn.Module):
    """Holds ``self.pose_net`` exactly like the reference's PoseNet* wrappers so that checkpoints
    whose keys start with ``pose_net.`` load with ``strict=True``.  The SMPL ParametricModel the
    reference attaches (a plain object, never in the state_dict; net_aagc.py:777) is not needed: the only thing
    ``forward_offline`` takes from it is the 24-entry parent table of the SMPL kinematic tree, which the
    reduced-global -> full-local kernel (``a3gc_reduced_to_full_local``) carries as a constant.
    """
    net_cls = None

    def __init__(self, input_size=12, rotsize=9, adjacency=None, device=torch.device("cpu"), n_hidden=256):
        super().__init__()
        self.rotsize = rotsize
        self.adjacency = adjacency
        self.pose_net = self.net_cls(input_size, rotsize, n_hidden, self.adjacency)
        self.rnn_state = None
        self.imu = None
        self.reset()

    def reset(self):
        self.rnn_state = None
        self.imu = None

    def forward(self, imu, rnn_state=None):
        global_reduced_pose, rnn_state = self.pose_net.forward(imu, rnn_state)
        return global_reduced_pose, rnn_state

    @torch.no_grad()
    def forward_offline(self, imu, rnn_state=None):
        global_reduced_pose, _ = self.forward(imu, rnn_state)
        if self.rotsize in (6, 9):
            from .pipeline import reduced_global_to_full_local
            return reduced_global_to_full_local(global_reduced_pose, self.rotsize), None      # net_aagc.py:825-828
        return global_reduced_pose, None


class PoseNet(_PoseNetBase):
    net_cls = AAGC_net


class PoseNet3(_PoseNetBase):
    net_cls = A3GC_net


class PoseNet_AGC(_PoseNetBase):
    net_cls = AGC_net


class PoseNet_GGRU(_PoseNetBase):
    net_cls = G_GRU_net


class pose_loss(torch.nn.Module):
    """net_aagc.py:1077-1087 (the reference forgets super().__init__(); this one is a proper Module)."""

    def __init__(self, loss_weight=None):
        super().__init__()
        self.loss_weight = loss_weight

    def forward(self, pred, targ):
        smpl_loss = torch.square(targ - pred)
        if self.loss_weight is not None:
            smpl_loss = smpl_loss * self.loss_weight
        return torch.mean(torch.sum(smpl_loss, -1, keepdim=False))
