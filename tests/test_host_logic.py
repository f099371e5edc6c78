"""CPU tests of the host-side mirror of the reference interface (no GPU needed)."""
import os

import pytest
import torch
import torch.multiprocessing as mp

import a3gc_ip_b200 as A
from conftest import load_golden
from oracle import net_oracle as O
from util import NET_CLS_NAMES, CELL_CLS_NAMES, trained_sd


@pytest.mark.parametrize("variant", O.VARIANTS)
@pytest.mark.parametrize("shape", [(12, 3, 256), (15, 3, 64), (15, 9, 128), (12, 3, 8)])
def test_state_dict_keys_shapes_order(variant, shape, nira):
    f0, out, hidden = shape
    net = getattr(A, NET_CLS_NAMES[variant])(f0, out, hidden, nira.float())
    got = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    assert got == O.net_param_shapes(variant, f0, out, hidden)
    frozen = {k for k, p in net.named_parameters() if not p.requires_grad}
    if variant == "AGC":
        assert frozen == {f"rnn{l}.directions.{d}.cell.adjacency" for l in (1, 2) for d in (0, 1)}
    elif variant == "GGRU":
        assert frozen == {f"rnn{l}.directions.{d}.cell.a" for l in (1, 2) for d in (0, 1)}
    else:
        assert not frozen


@pytest.mark.parametrize("name,cls,args", [("A3GC_model2", "PoseNet3", (15, 3, 64)), ("A3GC_model3", "PoseNet3", (15, 9, 128)),
                                           ("GGRU_model2", "PoseNet_GGRU", (15, 3, 64)), ("GGRU_model3", "PoseNet_GGRU", (15, 9, 128))])
def test_shipped_checkpoints_load_strict(name, cls, args, nira):
    ck = load_golden(os.path.join("weights", name + ".pt"))
    f0, rot, hidden = args
    net = getattr(A, cls)(input_size=f0, rotsize=rot, adjacency=nira.float(), n_hidden=hidden)
    res = net.load_state_dict(ck["state_dict"], strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in ck["state_dict"].items():
        assert torch.equal(net.state_dict()[k], v)


def test_parameters_never_alias(nira):
    tmpl = nira.float()
    before = tmpl.clone()
    for variant in O.VARIANTS:
        net = getattr(A, NET_CLS_NAMES[variant])(12, 3, 8, tmpl)
        ptrs = [p.data_ptr() for p in net.parameters()]
        assert len(set(ptrs)) == len(ptrs)
        assert tmpl.data_ptr() not in ptrs
    assert torch.equal(tmpl, before)        # the reference's G_GRU_cell ctor mutates the caller's template; ours must not


def test_init_matches_reference_scheme(nira):
    net = A.A3GC_net(12, 3, 16, nira.float())
    c = net.rnn1.directions[0].cell
    assert torch.equal(c.adjacency_i.data, nira.float().t())
    assert torch.count_nonzero(c.gcn_bias_i) == 0 and torch.count_nonzero(c.attention_bs) == 0
    bound = (6.0 / (16 + 32)) ** 0.5
    assert c.gcn_kernel_i.abs().max() <= bound and c.gcn_kernel_i.abs().max() > 0.5 * bound


def test_error_behaviour(nira):
    with pytest.raises(ValueError):
        A.A3GC_LSTM_cell(8, 8, nira.float(), activation_fn="relu")
    with pytest.raises(ValueError):
        A.AAGC(8, 8, nira.float(), activation_fn="sigmoid")
    with pytest.raises(AssertionError):
        A.A3GC_LSTM_cell(8, 8, torch.eye(14))
    with pytest.raises(AssertionError):
        A.G_GRU_cell(8, 8, torch.eye(24))
    net = A.A3GC_net(12, 3, 8, nira.float())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net.eval()(torch.zeros(1, 2, 15, 12))
    if torch.cuda.is_available():
        return
    with pytest.raises(ValueError):
        net.set_engine("cpu")


def test_missing_library_fails_loudly(monkeypatch):
    from a3gc_ip_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/liba3gc_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or eager fallback"):
        _lib.lib()


def test_pose_loss():
    g = load_golden("pose_loss.pt")
    assert torch.equal(A.pose_loss()(g["pred"], g["targ"]), g["loss"])


def test_shard_range_properties():
    for batch in (0, 1, 7, 8, 1024, 8191):
        for world in (1, 2, 3, 8):
            spans = [A.shard_range(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        A.shard_range(4, 2, 2)


def _gloo_worker(rank, world, port, batch, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = torch.arange(batch * 3, dtype=torch.float32).view(batch, 3)
        run = A.ShardedRunner()
        y_local = run.run(lambda t: t * 2 + 1, x)
        y = run.gather(y_local, batch)
        ret[rank] = bool(torch.equal(y, x * 2 + 1)) and y_local.shape[0] == A.shard_range(batch, world, rank)[1] - A.shard_range(batch, world, rank)[0]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [5, 8])
def test_sharded_runner_gloo_world2(batch):
    """N>1 path on CPU: two gloo ranks shard a ragged batch, compute independently, gather equals unsharded."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gloo_worker, args=(2, port, batch, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def _gloo_grad_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        lin = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        frozen = torch.nn.Parameter(torch.ones(4), requires_grad=False)
        params = list(lin.parameters()) + [frozen]
        x = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10.0
        t = torch.ones(8, 3)
        # every rank: local mean loss on its half of the batch -> all-reduced mean == gradient of the global-batch mean loss
        lo, hi = A.shard_range(8, world, rank)
        red = A.FlatGradAllReducer(params)
        loss = ((lin(x[lo:hi]) - t[lo:hi]) ** 2).sum(-1).mean()
        loss.backward()
        red.reduce()
        got = torch.cat([p.grad.flatten() for p in lin.parameters()])
        lin.zero_grad()
        ((lin(x) - t) ** 2).sum(-1).mean().backward()
        want = torch.cat([p.grad.flatten() for p in lin.parameters()])
        ret[rank] = bool(torch.allclose(got, want, rtol=1e-5, atol=1e-6)) and red.nbytes == 4 * sum(p.numel() for p in lin.parameters())
    finally:
        dist.destroy_process_group()


def test_flat_grad_allreduce_gloo_world2():
    """cfg-5 data parallelism on CPU: the flat-bucket all-reduce over two gloo ranks reproduces the global-batch gradient."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gloo_grad_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_training_loop_glue_follows_reference_control_flow(tmp_path):
    """fit_stage / checkpoint naming / resume discovery (train_a3gc_tp.py:164-187, 241-262) on a CPU stand-in model: the
    control flow is host logic and does not depend on the CUDA kernels."""
    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(4, 2)

        def forward(self, x, rnn_state=None):
            return self.lin(x), None

    torch.manual_seed(0)
    model, crit = Tiny(), A.pose_loss()
    xs = [torch.randn(8, 5, 4) for _ in range(3)]
    data = [(x, x[..., :2] * 2.0) for x in xs]
    vals = iter([5.0, 4.0, 4.5, 4.6, 4.7, 4.8, 4.9])       # improves twice, then never: stops once the counter exceeds patience = 2

    class FakeCrit:
        def forward(self, p, t):
            return crit.forward(p, t)

    import a3gc_ip_b200.train_loop as TL
    real_validate = TL.validate
    TL.validate = lambda m, c, b: next(vals)
    try:
        out = A.fit_stage(model, FakeCrit(), lambda: data, lambda: data, model_number=2, save_dir=str(tmp_path), lr=1e-2, patience=2,
                          max_epochs=50, log=lambda s: None)
    finally:
        TL.validate = real_validate
    assert [e for e, _, _ in out["history"]] == [0, 1, 2, 3, 4]            # epochs 2, 3, 4 do not improve: counter 3 > 2
    assert out["best_loss"] == 4.0 and out["checkpoint"].endswith("checkpoint_model2_pretrain_1.tar")
    assert abs(out["lr"] - 1e-2 * 0.8 ** 5) < 1e-12                        # ExponentialLR(0.8), one step per epoch
    ck = torch.load(out["checkpoint"])
    assert ck["epoch"] == 2 and set(ck["state_dict"]) == {"lin.weight", "lin.bias"}
    # resume discovery: highest epoch per model, 'pretrain' preferred when both kinds exist
    for name in ("checkpoint_model1_pretrain_3.tar", "checkpoint_model1_pretrain_12.tar", "checkpoint_model3_pretrain_7.tar",
                 "checkpoint_model1_finetuning_40.tar"):
        (tmp_path / name).write_bytes(b"")
    found = A.latest_checkpoints(str(tmp_path))
    assert os.path.basename(found[1]) == "checkpoint_model1_pretrain_12.tar"
    assert os.path.basename(found[2]) == "checkpoint_model2_pretrain_1.tar" and os.path.basename(found[3]) == "checkpoint_model3_pretrain_7.tar"
    assert A.checkpoint_name(3, 8, finetuning=True) == "checkpoint_model3_finetuning_8.tar"
    imu, a, b = torch.zeros(2, 3, 15, 12), torch.ones(2, 3, 15, 3), 2 * torch.ones(2, 3, 15, 3)
    x2, t2 = A.stage_inputs(2, imu, a, b, "leaf", "full", "smpl")
    assert x2.shape == (2, 3, 15, 15) and t2 == "full" and torch.equal(x2[..., 12:], a)
    assert torch.equal(A.stage_inputs(3, imu, a, b, "leaf", "full", "smpl")[0][..., 12:], b)


@pytest.mark.parametrize("total,window,hop", [(100, 40, 10), (100, 40, 40), (37, 40, 10), (41, 40, 7), (1000, 300, 30), (40, 40, 1)])
def test_window_plan_covers_every_frame_once(total, window, hop):
    """Windowed inference plan (TPPipeline.window_plan): the kept ranges tile [0, total) and lie inside their windows."""
    from a3gc_ip_b200.pipeline import TPPipeline
    starts, keep = TPPipeline.window_plan(total, window, hop)
    assert len(starts) == len(keep) and keep[0][0] == 0 and keep[-1][1] == total
    for (a, b), (c, d) in zip(keep, keep[1:]):
        assert b == c and a <= b
    for s0, (lo, hi) in zip(starts, keep):
        assert 0 <= s0 and s0 + min(window, total) <= total and lo >= s0 and hi <= s0 + window
    with pytest.raises(ValueError):
        TPPipeline.window_plan(total, window, window + 1)
