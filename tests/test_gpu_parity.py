"""Parity of the CUDA path (through the C ABI) with the reference's golden outputs and the CPU oracle.

Tolerances (SURVEY.md section 8d / BASELINE.json north_star):
  fp32 path: rel-L2 <= 1e-4 and max|err|/max|ref| <= 1e-4 against the reference's CPU fp32 forward;
  index / topology handling (prepare_input, stage concat): bit-exact.
"""
import os

import pytest
import torch

import a3gc_ip_b200 as A
from conftest import load_golden, rel_l2, max_rel
from oracle import net_oracle as O
from util import build_net, build_tp, case_sd, unflatten_h, flatten_h, CELL_CLS_NAMES, tp_state_dicts

pytestmark = pytest.mark.gpu
TOL = 1e-4
NET_CASES = load_golden("net_cases.pt")
CELL_CASES = load_golden("cell_cases.pt")


def engines_for(variant, f_in, hidden):
    """Every engine that claims support for this shape; SIMT always."""
    L = A.lib()
    out = ["simt"]
    if L.a3gc_layer_workspace_bytes(A._lib.VARIANT[variant], 8, 8, f_in, hidden, 2, 0, 2) > 0:
        out.append("tc")
    return out


def assert_close(got, want, tol=TOL, what=""):
    assert got.shape == want.shape, what
    assert torch.isfinite(got).all(), what
    r, m = rel_l2(got.cpu(), want), max_rel(got.cpu(), want)
    assert r <= tol and m <= tol, f"{what}: rel_l2={r:.3e} max_rel={m:.3e} (tol {tol})"


@pytest.mark.parametrize("case", NET_CASES, ids=[c["name"] for c in NET_CASES])
def test_net_matches_reference_golden(case, nira):
    sd = case_sd(case, nira)
    v, H = case["variant"], case["hidden"]
    engines = sorted(set(engines_for(v, H, H)) & set(engines_for(v, 2 * H, H))) or ["simt"]
    for eng in engines:
        net = build_net(v, case["f0"], case["out"], H, sd, nira, engine=eng)
        y, h = net(case["x"].cuda(), unflatten_h(v, case["h0"], "cuda"))
        assert_close(y, case["y"], what=f"{case['name']}[{eng}] y")
        for i, (a, b) in enumerate(zip(flatten_h(h), case["h_out"])):
            assert_close(a, b, what=f"{case['name']}[{eng}] state{i}")


@pytest.mark.parametrize("case", CELL_CASES, ids=[f'{c["variant"]}_{c["f_in"]}_{c["hidden"]}' for c in CELL_CASES])
def test_cell_matches_reference_golden(case, nira):
    v = case["variant"]
    cell = getattr(A, CELL_CLS_NAMES[v])(case["f_in"], case["hidden"], nira.float(), activation_fn="tanh")
    cell.load_state_dict(case["sd"], strict=True)
    cell = cell.cuda().eval().set_engine("simt")
    x, h, c = case["x"].cuda(), case["h"].cuda(), case["c"].cuda()
    if v == "GGRU":
        o, hn = cell(x, h)
        outs = [o, hn]
    else:
        o, (hn, cn) = cell(x, (h, c))
        outs = [o, hn, cn]
    for a, b in zip(outs, case["outs"]):
        assert_close(a, b, what=v)


@pytest.mark.parametrize("variant", O.VARIANTS)
def test_time_major_layers_match_oracle(variant, nira):
    """A3GC_LSTM / ReverseA3GC_LSTM style layers: time-major input, single direction."""
    H, F, T, B = 12, 20, 9, 5
    names = {"AAGC": ("AAGC_LSTM", "ReverseAAGC_LSTM"), "A3GC": ("A3GC_LSTM", "ReverseA3GC_LSTM"),
             "AGC": ("AGC_LSTM", "ReverseAGC_LSTM"), "GGRU": ("G_GRU", "ReverseG_GRU")}[variant]
    g = torch.Generator().manual_seed(3)
    x = torch.randn(T, B, 15, F, generator=g)
    h0, c0 = 0.3 * torch.randn(B, 15, H, generator=g), 0.3 * torch.randn(B, 15, H, generator=g)
    for rev, name in enumerate(names):
        layer = getattr(A, name)(F, H, nira.float(), activation_fn="tanh")
        for p_name, shape in O.cell_param_shapes(variant, F, H):
            obj = layer.cell
            for part in p_name.split("."):
                obj = getattr(obj, part)
            obj.data = 0.2 * torch.randn(shape, generator=g) + (nira.float().t() if shape == (15, 15) else 0)
        sd = {"l." + k: v for k, v in layer.state_dict().items()}
        state = h0 if variant == "GGRU" else (h0, c0)
        with torch.no_grad():
            want_y, want_s = O.layer_forward(variant, x, state, sd, "l.", reverse=bool(rev))
        layer = layer.cuda().eval().set_engine("simt")
        st = h0.cuda() if variant == "GGRU" else (h0.cuda(), c0.cuda())
        y, s = layer(x.cuda(), st)
        assert_close(y, want_y, what=name)
        for a, b in zip(flatten_h([s]), flatten_h([want_s])):
            assert_close(a, b, what=name + " state")


def test_aagc_graph_conv_matches_oracle(nira):
    g = torch.Generator().manual_seed(4)
    # gc_in with f_in known at compile time (12, 15) and generic (20), the generic kernel (64 -> 9), gc_out with 8 rows per warp
    # (f_out <= 4) and 4 rows per warp (f_out = 9 and 16)
    for f_in, f_out, act in ((12, 32, "linear"), (64, 9, "tanh"), (15, 256, "linear"), (512, 3, "linear"), (20, 32, "tanh"),
                             (256, 9, "linear"), (128, 16, "tanh"), (128, 3, "linear")):
        m = A.AAGC(f_in, f_out, nira.float(), activation_fn=act)
        m.gcn_bias.data = torch.randn(f_out, generator=g)
        m.adj.data += 0.1 * torch.randn(15, 15, generator=g)
        x = torch.randn(3, 7, 15, f_in, generator=g)                     # 21 frames: a ragged last group of 8 / 16
        sd = {"m." + k: v.clone() for k, v in m.state_dict().items()}
        want = O.aagc_forward(x, sd, "m.", act)
        assert_close(m.cuda().eval()(x.cuda()), want, what=f"AAGC {f_in}->{f_out}")


def test_prepare_input_and_concat_bit_exact():
    g = load_golden("prepare_input.pt")
    for tag, stats_name in (("nonorm", None), ("norm", "all_train_stats.pt"), ("norm_cda", "all_sym_train_stats.pt")):
        stats = None if stats_name is None else load_golden(stats_name)
        for ori, acc, want in zip(g["oris"], g["accs"], g["outs"][tag]):
            got = A.prepare_input(ori.cuda(), acc.cuda(), stats).unsqueeze(0).cpu()
            assert torch.equal(got, want), tag
    x = torch.randn(2, 3, 15, 12).cuda()
    p = torch.randn(2, 3, 15, 3).cuda()
    assert torch.equal(A.concat_stage_input(x, p), torch.cat((x, p), dim=-1))


@pytest.mark.parametrize("variant", ["A3GC", "GGRU"])
def test_tp_chain_cfg1_golden(variant, nira):
    """BASELINE cfg 1 (B=1, T=300, three stages) against the reference's own output."""
    g = load_golden("tp_cfg1.pt")[variant]
    pipe, _ = build_tp(variant, nira)
    x = O.synthetic_input(1, 300, seed=g["x_seed"]).cuda()
    y1, y2, y3 = pipe(x)
    for a, b, nm in ((y1, g["y1"], "y1"), (y2, g["y2"], "y2"), (y3, g["y3"], "y3")):
        assert_close(a, b, what=f"{variant} {nm}")


@pytest.mark.parametrize("variant", O.VARIANTS)
def test_edge_shapes(variant, nira):
    """Ragged and degenerate shapes: B not a multiple of the batch tile, T=1, B=1, empty batch."""
    sd = O.random_state_dict(variant, 12, 3, 24, nira, seed=9)
    net = build_net(variant, 12, 3, 24, sd, nira, engine="simt")
    for B, T in ((1, 1), (13, 3), (1, 17), (0, 4)):
        x = torch.randn(B, T, 15, 12, generator=torch.Generator().manual_seed(B * 100 + T))
        y, h = net(x.cuda())
        assert y.shape == (B, T, 15, 3)
        if B == 0:
            continue
        with torch.no_grad():
            want, want_h = O.net_forward(variant, x, sd)
        assert_close(y, want, what=f"{variant} B={B} T={T}")
        for a, b in zip(flatten_h(h), flatten_h(want_h)):
            assert_close(a, b, what=f"{variant} B={B} T={T} state")


def test_batch_independence_and_full_size_subset(nira):
    """BASELINE cfg 2 shape (B=1024, T=300 is run by bench.py; here B=256, T=300 keeps the test short):
    sequences are independent, so a strided subset run alone must reproduce the big-batch rows, and
    those rows must match the CPU oracle."""
    pipe, sds = build_tp("A3GC", nira)
    B, T = 256, 300
    x = O.synthetic_input(B, T, seed=1234)
    y3 = pipe(x.cuda())[2]
    idx = torch.arange(0, B, 37)
    y3_sub = pipe(x[idx].cuda())[2]
    assert_close(y3[idx.cuda()], y3_sub.cpu(), tol=2e-6, what="batch independence")
    with torch.no_grad():
        want = O.tp_forward("A3GC", x[idx[:3]], sds)[2]
    assert_close(y3[idx[:3].cuda()], want, what="big-batch rows vs oracle")


def test_linearity_of_graph_conv(nira):
    """Size-independent property at a large frame count: AAGC is affine in x."""
    m = A.AAGC(12, 64, nira.float()).cuda().eval()
    x1, x2 = torch.randn(64, 300, 15, 12).cuda(), torch.randn(64, 300, 15, 12).cuda()
    zero = m(torch.zeros_like(x1))
    lhs = m(x1 + 2 * x2) - zero
    rhs = (m(x1) - zero) + 2 * (m(x2) - zero)
    assert rel_l2(lhs.cpu(), rhs.cpu()) < 1e-5


def test_launch_counter_counts_kernels(nira):
    sd = O.random_state_dict("A3GC", 12, 3, 8, nira, seed=1)
    net = build_net("A3GC", 12, 3, 8, sd, nira, engine="simt")
    L = A.lib()
    L.a3gc_reset_launch_count()
    net(torch.zeros(2, 3, 15, 12).cuda())
    assert L.a3gc_launch_count() == 8      # in-GC, 2x(2 packs + layer), out-GC


@pytest.mark.parametrize("variant", ["AAGC", "AGC"])
def test_cfg3_bf16_tp_within_stated_bound(variant, nira):
    """BASELINE cfg 3 (AAGC-TP / AGC-TP, bf16 operands on the tensor-core engine, fp32 accumulate and state).
    Stated bound: rel-L2 <= 5e-3 and max-abs <= 2e-2 on O(1) outputs vs the CPU fp32 forward;
    the full 8192-sequence batch is the same kernel at 1024 sequences per GPU -- sequences are independent."""
    pipe, sds = build_tp(variant, nira, precision="bf16")
    B, T = 64, 300
    x = O.synthetic_input(B, T, seed=77)
    ys = pipe(x.cuda())
    idx = torch.tensor([0, 31, 63])
    with torch.no_grad():
        want = O.tp_forward(variant, x[idx], sds)
    for got, w, nm in zip(ys, want, ("y1", "y2", "y3")):
        g = got[idx.cuda()].cpu()
        r, m = rel_l2(g, w), float((g - w).abs().max())
        assert torch.isfinite(got).all()
        assert r <= 5e-3 and m <= 2e-2, f"{variant} bf16 {nm}: rel_l2={r:.3e} max_abs={m:.3e}"
    # and the fp32-parity path on the same inputs stays inside 1e-4
    pipe32, _ = build_tp(variant, nira)
    y32 = pipe32(x[idx].cuda())[2]
    assert_close(y32, want[2], what=f"{variant} fp32 TP")


def test_cfg4_ggru_long_sequence(nira):
    """BASELINE cfg 4 shape in time (G-GRU-TP, T=600): the recurrence must stay on the reference over 600 steps."""
    pipe, sds = build_tp("GGRU", nira)
    B, T = 48, 600
    x = O.synthetic_input(B, T, seed=4)
    y3 = pipe(x.cuda())[2]
    idx = torch.tensor([0, 47])
    with torch.no_grad():
        want = O.tp_forward("GGRU", x[idx], sds)[2]
    assert_close(y3[idx.cuda()], want, what="G-GRU T=600")


def test_ik_post_step_matches_reference_golden(nira):
    """forward_offline's reduced-global -> full-local pose kernel vs the reference's own IK (tests/golden/ik_cases.pt)."""
    g = load_golden("ik_cases.pt")
    y9 = A.reduced_global_to_full_local(g["x9"].cuda(), 9).cpu()
    y6 = A.reduced_global_to_full_local(g["x6"].cuda(), 6).cpu()
    assert y9.shape == g["y9"].shape and (y9 - g["y9"]).abs().max() <= 1e-6 * max(1.0, float(g["y9"].abs().max()))
    assert (y6 - g["y6"]).abs().max() <= 2e-6
    ident = torch.eye(3).expand(y9.shape[0], len(g["ignored"]), 3, 3)
    assert torch.equal(y9[:, g["ignored"]], ident)                       # index handling bit-exact
    # through the wrapper the reference's scripts call (evaluate_a3gc_tp.py:171): PoseNet3.forward_offline, rotsize 9
    net = A.PoseNet3(input_size=15, rotsize=9, adjacency=nira.float(), n_hidden=64).cuda().eval()
    x = torch.randn(2, 5, 15, 15).cuda()
    pose, _ = net.forward_offline(x)
    raw, _ = net.forward(x)
    want = O.reduced_global_to_full_local(raw.cpu().view(-1, 15, 3, 3), 9)
    assert pose.shape == (10, 24, 3, 3) and (pose.cpu() - want).abs().max() <= 1e-5
    with pytest.raises(ValueError):
        A.reduced_global_to_full_local(g["x9"].cuda(), 3)


@pytest.mark.parametrize("variant", O.VARIANTS)
def test_tensor_core_engine_ragged_batch_and_given_state(variant, nira):
    """tcgen05 engine edge cases: batch not a multiple of the 8-sequence tile (13, 1), T = 1, non-zero initial state,
    hidden 64 (one CTA per tile) and 128 (2-CTA clusters)."""
    for hidden in (64, 128):
        sd = O.random_state_dict(variant, 15, 3, hidden, nira, seed=31 + hidden)
        net = build_net(variant, 15, 3, hidden, sd, nira, engine="tc")
        for B, T in ((13, 3), (1, 1), (9, 2)):
            g = torch.Generator().manual_seed(B * 10 + T)
            x = torch.randn(B, T, 15, 15, generator=g)
            st = [0.3 * torch.randn(B, 15, hidden, generator=g) for _ in range(2 if variant == "GGRU" else 4)]
            h0 = unflatten_h(variant, st)
            y, h = net(x.cuda(), unflatten_h(variant, st, "cuda"))
            with torch.no_grad():
                want, want_h = O.net_forward(variant, x, sd, h0)
            assert_close(y, want, what=f"{variant} H={hidden} B={B} T={T}")
            for a, b in zip(flatten_h(h), flatten_h(want_h)):
                assert_close(a, b, what=f"{variant} H={hidden} B={B} T={T} state")


def test_concurrent_batch_chunks_and_host_path_match_single_stream(nira):
    """TPPipeline(streams=3): chunked concurrent execution and the host-buffer path reproduce the single-stream result."""
    pipe, _ = build_tp("A3GC", nira)
    x = O.synthetic_input(41, 20, seed=8)                 # ragged: 41 sequences -> chunks of 8-sequence tiles
    pipe.streams = 1
    want = [t.cpu() for t in pipe(x.cuda())]
    pipe.streams = 3
    got = [t.cpu() for t in pipe(x.cuda())]
    for a, b in zip(got, want):
        assert_close(a, b, tol=2e-6, what="chunked vs single stream")
    y = pipe.forward_host(x.pin_memory(), None, torch.device("cuda", 0))
    assert_close(y, want[2], tol=2e-6, what="host path (chunked)")
    pipe.streams = 1
    assert_close(pipe.forward_host(x, None, torch.device("cuda", 0)), want[2], tol=2e-6, what="host path (single stream)")


def test_sequential_macro_batches_match_single_pass(nira):
    """TPPipeline(max_frames=...): a call larger than the frame budget runs as sequential macro-batches of whole sequences
    (device and host-buffer paths); results equal the single-pass ones (sequences are independent)."""
    pipe, _ = build_tp("A3GC", nira)
    x = O.synthetic_input(29, 12, seed=9)
    pipe.streams = 2
    want = [t.cpu() for t in pipe(x.cuda())]
    pipe.max_frames = 12 * 16                             # 16 sequences per macro-batch -> 16 + 13
    assert pipe._macro_batch(29, 12) == 16
    got = [t.cpu() for t in pipe(x.cuda())]
    for a, b in zip(got, want):
        assert a.shape == b.shape
        assert_close(a, b, tol=2e-6, what="macro-batched vs single pass")
    assert_close(pipe.forward_host(x.pin_memory(), None, torch.device("cuda", 0)), want[2], tol=2e-6, what="macro-batched host path")
    pipe.max_frames = 1                                   # never below one 8-sequence tile
    assert pipe._macro_batch(29, 12) == 8
    assert_close(pipe(x.cuda())[2].cpu(), want[2], tol=2e-6, what="one tile per macro-batch")


def test_windowed_inference_matches_per_window_oracle(nira):
    """TPPipeline.forward_windowed: all sliding windows as one batch, each frame taken from its most central window.
    window >= T is the offline result; smaller windows equal the oracle run window by window and stitched by window_plan."""
    pipe, sds = build_tp("A3GC", nira)
    rec = O.synthetic_input(1, 44, seed=10)[0]                       # one recording [44, 15, 12]
    offline = pipe(rec.unsqueeze(0).cuda())[2][0].cpu()
    assert_close(pipe.forward_windowed(rec.cuda(), 64, 8).cpu(), offline, tol=2e-6, what="window >= T")
    window, hop = 16, 4
    starts, keep = pipe.window_plan(44, window, hop)
    assert len(starts) == 8 and keep[0][0] == 0 and keep[-1][1] == 44
    got = pipe.forward_windowed(rec.cuda(), window, hop).cpu()
    with torch.no_grad():
        for s0, (lo, hi) in zip(starts, keep):
            want = O.tp_forward("A3GC", rec[s0:s0 + window].unsqueeze(0), sds)[2][0]
            assert_close(got[lo:hi], want[lo - s0:hi - s0], what=f"window at {s0}")
