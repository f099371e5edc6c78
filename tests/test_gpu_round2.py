"""GPU parity tests added in round 2: the fused raw-input path (prepare_input + stage concat inside linear_in), the
tensor-core engine on the corners the first round left to the SIMT engine (cell / layer API, H = 256 clusters with ragged
batches and a caller-given state, wide net inputs), the A3GC bf16 bound, empty sequences, the NCCL gradient path and the
training-loop glue on the CUDA modules."""
import os
import subprocess
import sys

import pytest
import torch

import a3gc_ip_b200 as A
from conftest import load_golden, rel_l2, max_rel, ROOT
from oracle import net_oracle as O
from util import build_net, build_tp, unflatten_h, flatten_h, CELL_CLS_NAMES

pytestmark = pytest.mark.gpu
TOL = 1e-4


def assert_close(got, want, tol=TOL, what=""):
    assert got.shape == want.shape, what
    assert torch.isfinite(got).all(), what
    r, m = rel_l2(got.cpu(), want.cpu()), max_rel(got.cpu(), want.cpu())
    assert r <= tol and m <= tol, f"{what}: rel_l2={r:.3e} max_rel={m:.3e} (tol {tol})"


# ---------------------------------------------------------------------------------------------------------------
# fused prepare_input (evaluate_a3gc_tp.py:64-94) + stage concat (:168, :170)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("engine", ["tc", "simt"])
def test_fused_raw_input_is_bit_exact_vs_golden_prepare_input(engine, nira):
    """net.forward_raw(acc, ori) must equal net(prepare_input golden) BIT FOR BIT: the fused load performs the same
    arithmetic and the skipped zero nodes contribute exact zeros (index handling bit-exact)."""
    g = load_golden("prepare_input.pt")
    sd = O.random_state_dict("A3GC", 12, 3, 64, nira, seed=5)
    net = build_net("A3GC", 12, 3, 64, sd, nira, engine=engine)
    for tag, stats_name in (("nonorm", None), ("norm_cda", "all_sym_train_stats.pt")):
        stats = None if stats_name is None else load_golden(stats_name)
        for ori, acc, want_x in zip(g["oris"], g["accs"], g["outs"][tag]):
            y_ref, _ = net(want_x.cuda())                                     # the reference's own prepare_input output
            y_raw, _ = net.forward_raw(acc.float().unsqueeze(0).cuda(), ori.float().unsqueeze(0).cuda(), stats)
            assert torch.equal(y_raw, y_ref), f"{engine} {tag}"


def test_fused_raw_chain_matches_unfused_chain_and_oracle(nira):
    """Three stages from raw frames (pos of the previous stage concatenated inside linear_in) == prepare_input -> cat chain."""
    stats = load_golden("all_sym_train_stats.pt")
    pipe, sds = build_tp("A3GC", nira)
    pipe.stats = stats
    ori, acc = A.synthetic.synthetic_raw_imu(21, 24, seed=3, stats=stats)      # ragged batch
    for streams in (1, 3):
        pipe.streams = streams
        x = A.prepare_input(ori.cuda(), acc.cuda(), stats)
        want = pipe(x)
        got = pipe.forward_raw(ori.cuda(), acc.cuda())
        for a, b in zip(got, want):
            assert torch.equal(a, b), f"streams={streams}"
        y = pipe.forward_host_raw(ori.pin_memory(), acc.pin_memory(), None, torch.device("cuda", 0))
        assert torch.equal(y, want[2].cpu())
    with torch.no_grad():
        w = O.tp_forward("A3GC", O.prepare_input(ori[:2], acc[:2], stats), sds)[2]
    assert_close(got[2][:2], w, what="raw chain vs oracle")


def test_raw_input_argument_errors(nira):
    sd = O.random_state_dict("A3GC", 15, 3, 64, nira, seed=6)
    net = build_net("A3GC", 15, 3, 64, sd, nira)
    acc, ori = torch.zeros(2, 3, 18).cuda(), torch.zeros(2, 3, 54).cuda()
    with pytest.raises(RuntimeError):
        net.forward_raw(acc, ori)                       # units_in = 15 needs pos
    with pytest.raises(RuntimeError):
        net.forward_raw(acc.cpu(), ori)                 # no CPU path
    y, _ = net.forward_raw(acc, ori, None, torch.zeros(2, 3, 15, 3).cuda())
    assert y.shape == (2, 3, 15, 3)


# ---------------------------------------------------------------------------------------------------------------
# tensor-core engine corners
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", O.VARIANTS)
def test_tc_engine_h256_clusters_ragged_batch_and_given_state(variant, nira):
    """4-CTA clusters (H = 256) with B in {1, 13} and a caller-given non-zero (h0, c0)."""
    H = 256
    sd = O.random_state_dict(variant, 12, 3, H, nira, seed=41)
    net = build_net(variant, 12, 3, H, sd, nira, engine="tc")
    for B, T in ((1, 2), (13, 3)):
        g = torch.Generator().manual_seed(B * 7 + T)
        x = torch.randn(B, T, 15, 12, generator=g)
        st = [0.3 * torch.randn(B, 15, H, generator=g) for _ in range(2 if variant == "GGRU" else 4)]
        y, h = net(x.cuda(), unflatten_h(variant, st, "cuda"))
        with torch.no_grad():
            want, want_h = O.net_forward(variant, x, sd, unflatten_h(variant, st))
        assert_close(y, want, what=f"{variant} B={B} T={T}")
        for a, b in zip(flatten_h(h), flatten_h(want_h)):
            assert_close(a, b, what=f"{variant} B={B} T={T} state")


@pytest.mark.parametrize("variant", O.VARIANTS)
def test_cell_and_layer_api_on_tc_engine(variant, nira):
    """Cell (T = 1) and uni- / bidirectional layer (T = 9) calls on the tcgen05 engine (H = 64, F = 64)."""
    H, F, B = 64, 64, 5
    g = torch.Generator().manual_seed(12)
    cell = getattr(A, CELL_CLS_NAMES[variant])(F, H, nira.float(), activation_fn="tanh")
    for p in cell.parameters():
        if p.dim() == 1:
            p.data = 0.1 * torch.randn(p.shape, generator=g)
    sd = {"c." + k: v.clone() for k, v in cell.state_dict().items()}
    x1, h0, c0 = torch.randn(B, 15, F, generator=g), 0.3 * torch.randn(B, 15, H, generator=g), 0.3 * torch.randn(B, 15, H, generator=g)
    cell = cell.cuda().eval().set_engine("tc")
    with torch.no_grad():
        if variant == "GGRU":
            want = O.cell_ggru(x1, h0, sd, "c.")
            got = cell(x1.cuda(), h0.cuda())
            pairs = list(zip(got, want))
        else:
            wo, (wh, wc) = O.cell_lstm(variant, x1, (h0, c0), sd, "c.", activation="tanh")
            o, (hn, cn) = cell(x1.cuda(), (h0.cuda(), c0.cuda()))
            pairs = [(o, wo), (hn, wh), (cn, wc)]
    for a, b in pairs:
        assert_close(a, b, what=f"{variant} cell on tc")
    names = {"AAGC": ("AAGC_LSTM", "ReverseAAGC_LSTM", "BiAAGC_LSTM"), "A3GC": ("A3GC_LSTM", "ReverseA3GC_LSTM", "BiA3GC_LSTM"),
             "AGC": ("AGC_LSTM", "ReverseAGC_LSTM", "BiAGC_LSTM"), "GGRU": ("G_GRU", "ReverseG_GRU", "BiG_GRU")}[variant]
    T = 9
    x = torch.randn(T, B, 15, F, generator=g)
    for rev, name in enumerate(names[:2]):
        layer = getattr(A, name)(F, H, nira.float(), activation_fn="tanh")
        layer.cell.load_state_dict(cell.state_dict())
        lsd = {"l.cell." + k[2:]: v for k, v in sd.items()}
        state = h0 if variant == "GGRU" else (h0, c0)
        with torch.no_grad():
            want_y, want_s = O.layer_forward(variant, x, state, lsd, "l.", reverse=bool(rev))
        layer = layer.cuda().eval().set_engine("tc")
        st = h0.cuda() if variant == "GGRU" else (h0.cuda(), c0.cuda())
        y, s = layer(x.cuda(), st)
        assert_close(y, want_y, what=name + " on tc")
        for a, b in zip(flatten_h([s]), flatten_h([want_s])):
            assert_close(a, b, what=name + " state on tc")


def test_wide_net_input_falls_back_to_generic_linear_in(nira):
    """units_in > 32: engine=auto must not pick the fused linear_in -> image kernel; the reference accepts any width."""
    sd = O.random_state_dict("A3GC", 40, 3, 64, nira, seed=8)
    for engine in ("auto", "tc"):
        net = build_net("A3GC", 40, 3, 64, sd, nira, engine=engine)
        x = torch.randn(3, 4, 15, 40, generator=torch.Generator().manual_seed(2))
        y, _ = net(x.cuda())
        with torch.no_grad():
            want, _ = O.net_forward("A3GC", x, sd)
        assert_close(y, want, what=f"f0=40 [{engine}]")


def test_a3gc_bf16_within_stated_bound(nira):
    """A3GC-TP on the bf16 path: stated bound rel-L2 <= 5e-3, max-abs <= 2e-2, three stages, T = 300."""
    pipe, sds = build_tp("A3GC", nira, precision="bf16")
    x = O.synthetic_input(16, 300, seed=78)
    ys = pipe(x.cuda())
    idx = torch.tensor([0, 15])
    with torch.no_grad():
        want = O.tp_forward("A3GC", x[idx], sds)
    for got, w, nm in zip(ys, want, ("y1", "y2", "y3")):
        g = got[idx.cuda()].cpu()
        r, m = rel_l2(g, w), float((g - w).abs().max())
        assert torch.isfinite(got).all()
        assert r <= 5e-3 and m <= 2e-2, f"A3GC bf16 {nm}: rel_l2={r:.3e} max_abs={m:.3e}"


@pytest.mark.parametrize("variant", ["A3GC", "GGRU"])
def test_empty_sequence_returns_incoming_state(variant, nira):
    """T = 0: the reference's time loops never run and hand the state back (no uninitialised states)."""
    H = 64
    sd = O.random_state_dict(variant, 12, 3, H, nira, seed=3)
    net = build_net(variant, 12, 3, H, sd, nira)
    g = torch.Generator().manual_seed(1)
    st = [torch.randn(4, 15, H, generator=g) for _ in range(2 if variant == "GGRU" else 4)]
    y, h = net(torch.zeros(4, 0, 15, 12).cuda(), unflatten_h(variant, st, "cuda"))
    assert y.shape == (4, 0, 15, 3)
    for a, b in zip(flatten_h(h), st):
        assert torch.equal(a.cpu(), b)
    _, h = net(torch.zeros(4, 0, 15, 12).cuda())
    for a in flatten_h(h):
        assert torch.count_nonzero(a) == 0


def test_workspaces_are_not_copied_and_can_be_released(nira):
    import copy
    sd = O.random_state_dict("A3GC", 12, 3, 64, nira, seed=3)
    net = build_net("A3GC", 12, 3, 64, sd, nira)
    x = torch.randn(2, 3, 15, 12).cuda()
    y0, _ = net(x)
    assert net._ws.buf is not None
    twin = copy.deepcopy(net)
    assert twin._ws.buf is None
    net.release_workspaces()
    assert net._ws.buf is None
    assert torch.equal(net(x)[0], y0) and torch.equal(twin(x)[0], y0)


@pytest.mark.parametrize("variant", ["A3GC", "GGRU"])
def test_packed_weight_cache_is_exact_and_invalidates(variant, nira):
    """Opt-in packed-weight cache: same bits as per-call packing, refreshed after load_state_dict and after an
    optimizer-style in-place update (parameter version bump), shared safely between the concurrent streams of the pipeline."""
    sd = O.random_state_dict(variant, 12, 3, 64, nira, seed=21)
    net = build_net(variant, 12, 3, 64, sd, nira, engine="tc")
    x = torch.randn(9, 5, 15, 12, generator=torch.Generator().manual_seed(4)).cuda()
    want, _ = net(x)
    L = A.lib()
    net.cache_packed_weights(True)
    L.a3gc_reset_launch_count(); y1, _ = net(x); n_first = L.a3gc_launch_count()
    L.a3gc_reset_launch_count(); y2, _ = net(x); n_cached = L.a3gc_launch_count()
    assert torch.equal(y1, want) and torch.equal(y2, want)
    assert n_cached == 4 and n_first > n_cached                       # linear_in, rnn1, rnn2, linear_out: no pack kernels
    with torch.no_grad():                                              # what an optimizer step does
        for p in net.rnn2.parameters():
            p.mul_(1.01)
    fresh = build_net(variant, 12, 3, 64, {k: v.cpu() for k, v in net.state_dict().items()}, nira, engine="tc")
    assert torch.equal(net(x)[0], fresh(x)[0])
    net.load_state_dict(sd)
    assert torch.equal(net(x)[0], want)
    net.cache_packed_weights(False)
    assert torch.equal(net(x)[0], want)


def test_packed_weight_cache_with_concurrent_chunks(nira):
    pipe, _ = build_tp("A3GC", nira)
    x = O.synthetic_input(40, 12, seed=8).cuda()
    pipe.streams = 1
    want = pipe(x)[2]
    pipe.cache_packed_weights(True)
    pipe.streams = 4
    for _ in range(2):
        assert torch.equal(pipe(x)[2], want)


# ---------------------------------------------------------------------------------------------------------------
# training glue on the CUDA modules
# ---------------------------------------------------------------------------------------------------------------
def test_fit_stage_runs_on_cuda_modules_and_learns(tmp_path, nira):
    """fit_stage (train_a3gc_tp.py:241-262) driving the CUDA training path: loss decreases, a checkpoint with the reference's
    naming appears and reloads strict into the reference-shaped wrapper."""
    torch.manual_seed(0)
    model = A.PoseNet3(input_size=12, rotsize=3, adjacency=nira.float(), n_hidden=64).cuda()
    g = torch.Generator(device="cuda").manual_seed(1)
    ori = torch.randn(8, 12, 54, generator=g, device="cuda")
    acc = torch.randn(8, 12, 18, generator=g, device="cuda")
    full_pos = torch.randn(8, 12, 24, 3, generator=g, device="cuda")
    smpl = torch.randn(8, 12, 135, generator=g, device="cuda")
    sample = A.teacher_forced_sample(ori, acc, full_pos, smpl, None, generator=g)
    inputs, target = A.stage_inputs(1, *sample)
    assert inputs.shape == (8, 12, 15, 12) and target.shape == (8, 12, 45)
    # the reference's noise: std 0.025 on the teacher inputs only (datasets.py:54)
    noise = sample[2].reshape(8, 12, 15, 3) - full_pos[:, :, A.train_loop.SMPL_MAJOR_JOINTS]
    assert 0.015 < float(noise.std()) < 0.035 and torch.equal(sample[4].reshape(8, 12, 15, 3), full_pos[:, :, A.train_loop.SMPL_MAJOR_JOINTS])
    batches = lambda: [(inputs, target)]
    res = A.fit_stage(model, A.pose_loss(), batches, batches, 1, save_dir=str(tmp_path), lr=3e-3, patience=1, max_epochs=6, log=lambda s: None)
    hist = res["history"]
    assert hist[-1][1] < hist[0][1], hist
    assert res["checkpoint"] and os.path.basename(res["checkpoint"]).startswith("checkpoint_model1_pretrain_")
    ck = torch.load(res["checkpoint"])
    A.PoseNet3(input_size=12, rotsize=3, adjacency=nira.float(), n_hidden=64).load_state_dict(ck["state_dict"], strict=True)


def test_second_backward_raises_clear_error(nira):
    net = A.A3GC_net(12, 3, 64, nira.float(), 0.0, 0.0, 0.0).cuda().train()
    y, _ = net(torch.randn(2, 3, 15, 12).cuda())
    loss = y.square().sum()
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="ran twice"):
        loss.backward()


NCCL_CHILD = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["A3GC_ROOT"])
import a3gc_ip_b200 as A
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
nira = A.synthetic.load_nira()
torch.manual_seed(0)
net = A.A3GC_net(12, 3, 64, nira, 0.0, 0.0, 0.0).to(dev).train()
g = torch.Generator().manual_seed(5)
B = 8
x = torch.randn(B, 6, 15, 12, generator=g); t = torch.randn(B, 6, 45, generator=g)
crit = A.pose_loss()
# global-batch gradient on this rank alone
y, _ = net(x.to(dev)); loss = crit.forward(y.view(B, 6, 45), t.to(dev)); net.zero_grad(); loss.backward()
want = [p.grad.clone() for p in net.parameters()]
for overlap in (True, False):
    red = A.FlatGradAllReducer.for_net(net, overlap=overlap)
    lo, hi = A.shard_range(B, world, rank)
    y, _ = net(x[lo:hi].to(dev)); loss = crit.forward(y.view(hi - lo, 6, 45), t[lo:hi].to(dev))
    red.zero_grad(); loss.backward(); red.reduce(hi - lo)
    torch.cuda.synchronize()
    err = max(float((p.grad - w).norm() / w.norm().clamp_min(1e-12)) for p, w in zip(net.parameters(), want))
    assert err < 1e-4, (overlap, err)
    assert red._back_issued == overlap
    for p in net.parameters(): p.grad = None
if rank == 0: print("NCCL_GRAD_OK")
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_all_reduced_gradients_equal_global_batch_gradients(tmp_path):
    """Two NCCL ranks, each with half of the batch: the all-reduced (overlapped and serial) gradients equal the gradients of
    the global batch computed on one rank."""
    script = tmp_path / "child.py"
    script.write_text(NCCL_CHILD)
    env = dict(os.environ, A3GC_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "NCCL_GRAD_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
