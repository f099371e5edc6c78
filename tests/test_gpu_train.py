"""Training path (BASELINE cfg 5): forward in train() mode + BPTT backward through the C ABI vs autograd through the
CPU oracle (same weights, same inputs, dropout = 0 -- the reference's masks come from the TorchScript RNG and cannot
be reproduced).  Tolerance: rel-L2 <= 1e-4 on the loss, the outputs and every parameter / input gradient."""
import pytest
import torch

import a3gc_ip_b200 as A
from conftest import rel_l2
from oracle import net_oracle as O
from util import NET_CLS_NAMES, flatten_h

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _oracle_grads(variant, x, sd, target, h=None):
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    y, hout = O.net_forward(variant, xr, sd, h)
    loss = torch.mean(torch.sum(torch.square(target - y.reshape(target.shape)), -1))       # pose_loss, net_aagc.py:1081-1087
    loss.backward()
    return loss.detach(), y.detach(), xr.grad, {k: v.grad for k, v in sd.items()}


@pytest.mark.parametrize("variant,hidden,B,T", [("A3GC", 24, 5, 7), ("AAGC", 24, 3, 6), ("AGC", 24, 4, 5),      # CUDA-core forward
                                                ("A3GC", 64, 3, 9), ("A3GC", 128, 2, 4), ("AAGC", 64, 9, 3),     # tcgen05 forward
                                                ("AGC", 128, 11, 5), ("A3GC", 256, 2, 3),
                                                ("GGRU", 24, 5, 6), ("GGRU", 64, 9, 4)])                         # graph-GRU (CUDA-core)
def test_net_train_step_matches_oracle_autograd(variant, hidden, B, T, nira):
    f0, out = 15, 9
    sd = O.random_state_dict(variant, f0, out, hidden, nira, seed=21)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, T, 15, f0, generator=g)
    target = torch.randn(B, T, 15 * out, generator=g)
    want_loss, want_y, want_dx, want_g = _oracle_grads(variant, x, sd, target)

    net = getattr(A, NET_CLS_NAMES[variant])(f0, out, hidden, nira.float(), linear_dropout=0.0, dropout=0.0, recurrent_dropout=0.0)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train()
    xc = x.cuda().requires_grad_(True)
    y, _ = net(xc)
    loss = A.pose_loss()(y.view(B, T, 15 * out), target.cuda())
    loss.backward()
    assert rel_l2(y.detach().cpu(), want_y) <= TOL
    assert abs(float(loss) - float(want_loss)) <= TOL * abs(float(want_loss))
    assert rel_l2(xc.grad.cpu(), want_dx) <= TOL, f"dx {rel_l2(xc.grad.cpu(), want_dx):.3e}"
    for name, p in net.named_parameters():
        w = want_g[name]
        if not p.requires_grad:
            assert p.grad is None
            continue
        assert p.grad is not None, name
        r = rel_l2(p.grad.cpu(), w)
        assert r <= TOL, f"{variant} H={hidden} grad {name}: rel_l2={r:.3e}"


def test_bilayer_train_with_initial_state_and_state_grads(nira):
    """A Bi layer on its own: non-zero (h0, c0), gradients flowing in through y AND the final states (as rnn2 -> rnn1 do)."""
    variant, F, H, B, T = "A3GC", 20, 16, 4, 6
    g = torch.Generator().manual_seed(9)
    layer = A.BiA3GC_LSTM(F, H, nira.float(), activation_fn="tanh")
    for p in layer.parameters():
        p.data = 0.2 * torch.randn(p.shape, generator=g) + (nira.float().t() if tuple(p.shape) == (15, 15) else 0)
    sd = {"l." + k: v.detach().clone() for k, v in layer.state_dict().items()}
    x = torch.randn(B, T, 15, F, generator=g)
    st = [tuple(0.3 * torch.randn(B, 15, H, generator=g) for _ in range(2)) for _ in range(2)]
    wy = torch.randn(B, T, 15, 2 * H, generator=g)
    ws = [tuple(torch.randn(B, 15, H, generator=g) for _ in range(2)) for _ in range(2)]

    def scalar(y, states):
        tot = (y * wy.to(y.device)).sum()
        for (h, c), (a, b) in zip(states, ws):
            tot = tot + (h * a.to(y.device)).sum() + (c * b.to(y.device)).sum()
        return tot

    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    str_ = [tuple(t.clone().requires_grad_(True) for t in s) for s in st]
    y, so = O.bi_layer_forward(variant, xr, str_, sdr, "l.")
    scalar(y, so).backward()

    layer = layer.cuda().train()
    xc = x.cuda().requires_grad_(True)
    stc = [tuple(t.cuda().requires_grad_(True) for t in s) for s in st]
    yc, soc = layer(xc, stc)
    scalar(yc, soc).backward()
    assert rel_l2(yc.detach().cpu(), y.detach()) <= TOL
    assert rel_l2(xc.grad.cpu(), xr.grad) <= TOL
    for a, b in zip(flatten_h(stc), flatten_h(str_)):
        assert rel_l2(a.grad.cpu(), b.grad) <= TOL
    for name, p in layer.named_parameters():
        r = rel_l2(p.grad.cpu(), sdr["l." + name].grad)
        assert r <= TOL, f"grad {name}: rel_l2={r:.3e}"


@pytest.mark.parametrize("hidden", [16, 64])
def test_train_mode_dropout_runs_and_is_stochastic(hidden, nira):
    net = A.A3GC_net(12, 3, hidden, nira.float()).cuda().train()      # reference defaults: p = 0.2 / 0.3 / 0.3
    x = O.synthetic_input(3, 5, seed=1).cuda()
    y1, _ = net(x)
    y2, _ = net(x)
    assert torch.isfinite(y1).all() and not torch.equal(y1, y2)
    y1.square().mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters() if p.requires_grad)


def test_recurrent_dropout_mask_semantics_tc_vs_simt(nira):
    """Same explicit recurrent-dropout mask through the tcgen05 and the CUDA-core training forwards (and backwards):
    identical outputs and gradients, i.e. both engines apply the mask to the h that enters the gates only."""
    from a3gc_ip_b200 import training as TR
    torch.manual_seed(3)
    F, H, B, T = 32, 64, 5, 6
    layer = A.BiA3GC_LSTM(F, H, nira.float(), activation_fn="tanh").cuda().train()
    x = torch.randn(B, T, 15, F, device="cuda")
    st = [tuple(0.2 * torch.randn(B, 15, H, device="cuda") for _ in range(2)) for _ in range(2)]
    hmask = (torch.rand(2, B, T, 15, H, device="cuda") >= 0.3).float() / 0.7
    outs = {}
    for eng in ("simt", "tc"):
        flat = []
        for s_ in st:
            flat += [s_[0], s_[1]]
        for d in layer.directions:
            flat += [getattr(d.cell, n) for n in TR.LSTM_PARAM_NAMES["A3GC"]]
        for p_ in layer.parameters():
            p_.grad = None
        xr = x.clone().requires_grad_(True)
        res = TR._LayerTrainFn.apply(("A3GC", 2, (0, 1), "tanh", A._lib.Workspace(), eng), xr, hmask, *flat)
        (res[0].square().sum() + sum(r.sum() for r in res[1:])).backward()
        outs[eng] = [res[0].detach().cpu(), xr.grad.cpu()] + [p_.grad.detach().cpu().clone() for p_ in layer.parameters()]
    for a, b in zip(outs["simt"], outs["tc"]):
        assert rel_l2(b, a) <= TOL


@pytest.mark.parametrize("variant,H,B", [("A3GC", 64, 5), ("AAGC", 128, 3), ("AGC", 256, 3)])
def test_blocked_and_plain_backward_chains_agree(variant, H, B, nira, monkeypatch):
    """The blocked reverse-time chain (H in {64,128,256}; its two weight contractions as 3xTF32 mma.sync products) against the
    one-(sequence, unit)-per-thread fp32 chain on the same tape, with a recurrent-dropout mask and a ragged batch.  Bound 3e-5:
    the hi/lo TF32 split drops the lo x lo term (2^-22 per product); both chains are within 1e-4 of the oracle separately."""
    from a3gc_ip_b200 import training as TR
    torch.manual_seed(5)
    F, T = 32, 5
    layer = getattr(A, f"Bi{variant}_LSTM")(F, H, nira.float(), activation_fn="tanh").cuda().train()
    x = torch.randn(B, T, 15, F, device="cuda")
    st = [tuple(0.2 * torch.randn(B, 15, H, device="cuda") for _ in range(2)) for _ in range(2)]
    hmask = (torch.rand(2, B, T, 15, H, device="cuda") >= 0.3).float() / 0.7
    outs = {}
    for blk in ("1", "0"):
        monkeypatch.setenv("A3GC_BWD_BLK", blk)
        flat = []
        for s_ in st:
            flat += [s_[0].clone().requires_grad_(True), s_[1].clone().requires_grad_(True)]
        for d in layer.directions:
            flat += [getattr(d.cell, n) for n in TR.LSTM_PARAM_NAMES[variant]]
        for p_ in layer.parameters():
            p_.grad = None
        xr = x.clone().requires_grad_(True)
        res = TR._LayerTrainFn.apply((variant, 2, (0, 1), "tanh", A._lib.Workspace(), "tc"), xr, hmask, *flat)
        (res[0].square().sum() + sum(r.sum() for r in res[1:])).backward()
        outs[blk] = ([xr.grad.cpu()] + [f.grad.cpu() for f in flat[:4]]
                     + [p_.grad.detach().cpu().clone() for p_ in layer.parameters() if p_.grad is not None])
    assert len(outs["1"]) == len(outs["0"]) > 10
    for i, (a, b) in enumerate(zip(outs["0"], outs["1"])):
        assert rel_l2(b, a) <= 3e-5, f"gradient {i}: rel_l2={rel_l2(b, a):.3e}"


def test_tf32_split_and_hprev_operand_builders(monkeypatch):
    """a3gc_train_split_tf32: hi is exactly representable in TF32 (13 low significand bits zero), hi + lo == x bit for bit;
    a3gc_train_hprev_split: the shifted, masked h_prev operand equals the torch construction (net_aagc.py:181-182)."""
    from a3gc_ip_b200 import training as TR
    monkeypatch.setenv("A3GC_TRAIN_GEMM", "tf32x3")
    torch.manual_seed(11)
    x = (torch.randn(1027, 33, device="cuda") * torch.logspace(-20, 6, 33, device="cuda")).contiguous()
    sp = TR._split(x)
    assert int((sp.hi.view(torch.int32) & 0x1FFF).abs().max()) == 0
    assert torch.equal(sp.hi + sp.lo, x)
    assert float(((sp.lo.abs() > 0) & (sp.lo.abs() > sp.hi.abs() * 2.0 ** -10)).sum()) == 0      # |lo| <= half an ulp of TF32
    B, T, H = 5, 7, 64
    hp = torch.randn(B, T, 15, H, device="cuda")
    h0 = torch.randn(B, 15, H, device="cuda")
    mask = (torch.rand(B, T, 15, H, device="cuda") >= 0.3).float() / 0.7
    for reverse in (0, 1):
        for m in (None, mask):
            for z in (h0, None):
                got = TR._hprev_split(hp, z, m, reverse)
                first = z if z is not None else torch.zeros_like(h0)
                want = torch.cat((hp[:, 1:], first.unsqueeze(1)), dim=1) if reverse else torch.cat((first.unsqueeze(1), hp[:, :-1]), dim=1)
                if m is not None:
                    want = want * m
                assert torch.equal((got.hi + got.lo).reshape(B, T, 15, H), want)
                assert int((got.hi.view(torch.int32) & 0x1FFF).abs().max()) == 0


def test_mixed_operand_builders_and_gemms():
    """a3gc_train_split_mixed / a3gc_train_hprev_split_mixed (ABI 3): hi is TF32-exact, hi16 = bf16(hi), lo16 = bf16(x - hi), written
    at a column offset of a wider row-major buffer (x and h_prev side by side); the mixed products (TF32 hi x hi + two bf16
    correction GEMMs) agree with an fp64 product to fp32-level accuracy."""
    from a3gc_ip_b200 import training as TR
    torch.manual_seed(12)
    R, F, H = 4 * 6 * 15, 24, 64
    x = torch.randn(R, F, device="cuda") * torch.logspace(-6, 3, F, device="cuda")
    B, T = 4, 6
    hp = torch.randn(B, T, 15, H, device="cuda")
    h0 = torch.randn(B, 15, H, device="cuda")
    mask = (torch.rand(B, T, 15, H, device="cuda") >= 0.3).float() / 0.7
    S = TR._mixed_buffers(R, F + H, "cuda")
    TR._split_into(S, x, 0)
    TR._hprev_split(hp, h0, mask, 1, into=S, col0=F)
    want = torch.cat((x, (torch.cat((hp[:, 1:], h0.unsqueeze(1)), dim=1) * mask).reshape(R, H)), dim=1)
    assert int((S.hi.view(torch.int32) & 0x1FFF).abs().max()) == 0
    assert torch.equal(S.hi16, S.hi.bfloat16())
    assert torch.equal(S.lo16, (want - S.hi).bfloat16())
    assert float(((want - S.hi).abs() > S.hi.abs() * 2.0 ** -10).sum()) == 0
    a = torch.randn(R, 96, device="cuda")
    got = TR._mm_tn(TR._split(a), S)
    ref = (a.double().t() @ want.double()).float()
    assert rel_l2(got.cpu(), ref.cpu()) <= 2e-6
    w = torch.randn(F + H, 40, device="cuda")
    out = torch.zeros(R, 40, device="cuda")
    TR._addmm_nn(out, S, TR._split(w))
    assert rel_l2(out.cpu(), (want.double() @ w.double()).float().cpu()) <= 2e-6
    odd = torch.randn(R, 15, device="cuda")                       # a width the vectorised builder does not take: three-pass form
    got = TR._mm_tn(TR._split(a), TR._split(odd))
    assert rel_l2(got.cpu(), (a.double().t() @ odd.double()).float().cpu()) <= 2e-6


def test_fused_adjacency_gradient_kernel():
    """a3gc_train_adjacency_grad against the bmm + sum it replaces (training.py), fp64 reference; deterministic across calls."""
    from a3gc_ip_b200 import _lib
    torch.manual_seed(13)
    R, H = 333, 64
    dz = torch.randn(R, 4, H, 16, device="cuda")
    u = torch.randn(R, 4, H, 16, device="cuda")
    want = torch.einsum("rgjm,rgjn->gmn", dz.double(), u.double()).float()
    outs = []
    for nblk in (7, 64):
        part, dP = torch.empty(nblk, 1024, device="cuda"), torch.empty(4, 16, 16, device="cuda")
        rc = A.lib().a3gc_train_adjacency_grad(dz.data_ptr(), u.data_ptr(), R, H, part.data_ptr(), nblk, dP.data_ptr(), _lib.stream_ptr(dz.device))
        _lib.check(rc, "a3gc_train_adjacency_grad")
        assert rel_l2(dP.cpu(), want.cpu()) <= 2e-6
        outs.append(dP.clone())
    part, dP = torch.empty(64, 1024, device="cuda"), torch.empty(4, 16, 16, device="cuda")
    A.lib().a3gc_train_adjacency_grad(dz.data_ptr(), u.data_ptr(), R, H, part.data_ptr(), 64, dP.data_ptr(), _lib.stream_ptr(dz.device))
    assert torch.equal(dP, outs[1])
