"""BASELINE.json's full sizes through size-independent properties (the oracle only runs a few sequences in seconds):
sequences are independent, so rows of the big batch must (a) equal the same rows run as a small batch and (b) match the CPU
oracle; outputs must be finite everywhere."""
import pytest
import torch

import a3gc_ip_b200 as A
from conftest import rel_l2, max_rel
from oracle import net_oracle as O
from util import build_tp

pytestmark = pytest.mark.gpu


def _check_rows(pipe, sds, variant, x, idx, tol, tol_abs=None):
    ys = pipe(x.cuda())
    for y in ys:
        assert torch.isfinite(y).all()
    sub = pipe(x[idx].cuda())
    for a, b in zip(ys, sub):
        assert rel_l2(a[idx.cuda()].cpu(), b.cpu()) <= 2e-6            # batch independence (incl. chunked multi-stream execution)
    with torch.no_grad():
        want = O.tp_forward(variant, x[idx[:3]], sds)
    for a, w in zip(ys, want):
        g = a[idx[:3].cuda()].cpu()
        if tol_abs is None:
            assert rel_l2(g, w) <= tol and max_rel(g, w) <= tol
        else:
            assert rel_l2(g, w) <= tol and float((g - w).abs().max()) <= tol_abs


def test_cfg2_full_size_a3gc_tp_fp32(nira):
    """cfg 2: A3GC-TP fp32, B = 1024 x T = 300 on one GPU."""
    pipe, sds = build_tp("A3GC", nira)
    pipe.streams = 4
    x = O.synthetic_input(1024, 300, seed=1234)
    _check_rows(pipe, sds, "A3GC", x, torch.tensor([0, 511, 1023, 77, 640]), 1e-4)


@pytest.mark.parametrize("variant", ["AAGC", "AGC"])
def test_cfg3_per_gpu_share_bf16(variant, nira):
    """cfg 3: AAGC-TP / AGC-TP bf16, the 1024-sequence share one GPU owns of the 8192-sequence batch, T = 300."""
    pipe, sds = build_tp(variant, nira, precision="bf16")
    x = O.synthetic_input(1024, 300, seed=1234 + 3)
    _check_rows(pipe, sds, variant, x, torch.tensor([5, 1000, 512]), 5e-3, tol_abs=2e-2)


def test_cfg4_per_gpu_share_ggru_long(nira):
    """cfg 4: G-GRU-TP, T = 600, the 512-sequence share of one of 8 GPUs (4096 / 8)."""
    pipe, sds = build_tp("GGRU", nira)
    x = O.synthetic_input(512, 600, seed=1234 + 4)
    _check_rows(pipe, sds, "GGRU", x, torch.tensor([0, 300, 511]), 1e-4)


def test_cfg5_full_size_training_step_runs_and_learns(nira):
    """cfg 5 shape (B = 256 x T = 200, stage 3 = the H = 128 net): a few optimisation steps on a fixed batch reduce the loss, all
    gradients finite (gradient VALUES are checked against oracle autograd at small sizes in test_gpu_train.py)."""
    torch.manual_seed(0)
    net = A.A3GC_net(15, 9, 128, nira.float()).cuda().train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    crit = A.pose_loss()
    x = torch.randn(256, 200, 15, 15, device="cuda")
    tgt = 0.1 * torch.randn(256, 200, 135, device="cuda")
    losses = [float(A.train_step(net, crit, opt, x, tgt)) for _ in range(4)]
    assert all(torch.isfinite(p.grad).all() for p in net.parameters() if p.grad is not None)
    assert losses[-1] < losses[0]
