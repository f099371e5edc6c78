"""CPU tests (no GPU) of the round-2 host logic: the flat-bucket gradient reducer with its two sub-buckets, batch-weighted
mean and ragged shards, data-parallel fit_stage decisions, the synthetic-workload module, the staged reference and the
reference arm of bench.py."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch
import torch.multiprocessing as mp

import a3gc_ip_b200 as A
from conftest import ROOT


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


class _Net(torch.nn.Module):
    """Parameter names of the nets (linear_in, rnn1 | rnn2, linear_out) so that for_net() finds the bucket cut."""

    def __init__(self):
        super().__init__()
        self.linear_in = torch.nn.Linear(6, 5)
        self.rnn1 = torch.nn.Linear(5, 5)
        self.rnn2 = torch.nn.Linear(5, 4)
        self.linear_out = torch.nn.Linear(4, 3)

    def forward(self, x, h=None):
        return self.linear_out(torch.tanh(self.rnn2(torch.tanh(self.rnn1(torch.tanh(self.linear_in(x))))))), None


def _reducer_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = _Net()
        x = torch.arange(7 * 6, dtype=torch.float32).view(7, 6) / 10.0
        t = torch.ones(7, 3)
        net.zero_grad()
        ((net(x)[0] - t) ** 2).sum(-1).mean().backward()
        want = torch.cat([p.grad.flatten() for p in net.parameters()]).clone()
        ok = True
        # ragged shards (4 + 3 sequences): the serial form weights every rank by its batch size
        bounds = [(0, 4), (4, 7)][rank]
        red = A.FlatGradAllReducer.for_net(net, overlap=False)
        ok &= red.split == 4                                     # linear_in.{w,b}, rnn1.{w,b} | rnn2, linear_out
        red.zero_grad()
        ok &= all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(red.params, red.views))      # grads ARE the bucket
        ((net(x[bounds[0]:bounds[1]])[0] - t[bounds[0]:bounds[1]]) ** 2).sum(-1).mean().backward()
        red.reduce(bounds[1] - bounds[0])
        ok &= bool(torch.allclose(red.bucket, want, rtol=1e-5, atol=1e-6))
        # equal shards (rank 1 drops one sequence: compare with the 6-sequence global batch), overlapped form: the back half
        # goes out from the post-accumulate hook while "rnn1" is still in backward
        x6, t6 = x[:6], t[:6]
        net.zero_grad(set_to_none=True)
        ((net(x6)[0] - t6) ** 2).sum(-1).mean().backward()
        want6 = torch.cat([p.grad.flatten() for p in net.parameters()]).clone()
        red2 = A.FlatGradAllReducer.for_net(net, overlap=True)
        red2.zero_grad()
        lo, hi = 3 * rank, 3 * rank + 3
        ((net(x6[lo:hi])[0] - t6[lo:hi]) ** 2).sum(-1).mean().backward()
        ok &= red2._back_issued
        red2.reduce(3)
        ok &= bool(torch.allclose(red2.bucket, want6, rtol=1e-5, atol=1e-6))
        # the overlapped form refuses ragged shards instead of returning a wrong mean (checked on its first step)
        for p_ in net.parameters():
            p_.grad = None
        red2.close()
        red2 = A.FlatGradAllReducer.for_net(net, overlap=True)
        red2.zero_grad()
        ((net(x[bounds[0]:bounds[1]])[0] - t[bounds[0]:bounds[1]]) ** 2).sum(-1).mean().backward()
        try:
            red2.reduce(bounds[1] - bounds[0])
            ok = False
        except RuntimeError:
            pass
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


def test_flat_grad_reducer_buckets_weights_and_overlap_gloo_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_reducer_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def _fit_worker(rank, world, port, tmp, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = _Net()
        g = torch.Generator().manual_seed(7)
        x = torch.randn(12, 6, generator=g)
        t = torch.randn(12, 3, generator=g)
        lo, hi = (0, 7) if rank == 0 else (7, 12)                  # ragged shards, very different local validation losses
        train = lambda: [(x[lo:hi], t[lo:hi])]
        valid = lambda: [(x[lo:hi] * (1 + 3 * rank), t[lo:hi])]
        out = A.fit_stage(net, A.pose_loss(), train, valid, 1, save_dir=tmp, lr=1e-2, patience=1, max_epochs=8, data_parallel=True,
                          log=lambda s: None)
        w = torch.cat([p.detach().flatten() for p in net.parameters()])
        ret[rank] = (out["best_loss"], len(out["history"]), [h[2] for h in out["history"]], w.tolist(), out["checkpoint"])
    finally:
        dist.destroy_process_group()


def test_fit_stage_data_parallel_ranks_agree_gloo_world2(tmp_path):
    """every rank must take the same improvement / early-stop decisions (global validation loss), end with the same
    weights, and only rank 0 writes the checkpoint."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_fit_worker, args=(2, _free_port(), str(tmp_path), ret), nprocs=2, join=True)
    a, b = ret[0], ret[1]
    assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2]
    assert torch.allclose(torch.tensor(a[3]), torch.tensor(b[3]), rtol=1e-6, atol=1e-7)
    assert a[4] == b[4] and os.path.exists(a[4])
    assert len([f for f in os.listdir(tmp_path) if f.startswith("checkpoint_model1")]) >= 1


def test_synthetic_workload_module():
    S = A.synthetic
    stats = S.load_stats()
    ori, acc = S.synthetic_raw_imu(3, 5, seed=1, stats=stats)
    assert ori.shape == (3, 5, 54) and acc.shape == (3, 5, 18)
    # normalising the synthetic raw frames gives unit-normal channels: (v - mean) / std is what prepare_input applies
    z = (ori - stats["ori"]["mean_channel"].float()) / stats["ori"]["std_channel"].float()
    o2, _ = S.synthetic_raw_imu(3, 5, seed=1, stats=None)
    assert torch.allclose(z, o2, atol=1e-4)
    x = S.synthetic_input(2, 4, seed=3)
    assert x.shape == (2, 4, 15, 12) and torch.count_nonzero(x[:, :, [0, 1, 2, 5, 6, 7, 8, 9, 11, 12]]) == 0
    nira = S.load_nira()
    sds = S.tp_state_dicts("A3GC", nira)
    assert [sd["linear_in.gcn_kernel"].shape for sd in sds] == [(256, 12), (64, 15), (128, 15)]
    assert torch.equal(S.random_state_dict("AGC", 12, 3, 8, nira, 5)["linear_in.gcn_kernel"], S.random_state_dict("AGC", 12, 3, 8, nira, 5)["linear_in.gcn_kernel"])
    for sd, (f0, o, h) in zip(sds, S.TP_SHAPES):
        A.A3GC_net(f0, o, h, nira).load_state_dict(sd, strict=True)


def test_staged_reference_is_byte_identical_and_matches_the_oracle():
    from oracle import build_ref, net_oracle as O
    if not build_ref.available():
        if not os.path.isdir(build_ref.REF):
            pytest.skip("oracle/_ref was not staged and /root/reference is absent")
        build_ref.stage()
    if os.path.isdir(build_ref.REF):
        assert build_ref.stage(check_only=True)
    S = A.synthetic
    sds = S.tp_state_dicts("A3GC", S.load_nira())
    nets = build_ref.ref_tp_nets("A3GC", sds)
    x = S.synthetic_input(2, 6, seed=5)
    y = build_ref.ref_tp_forward(nets, x)
    with torch.no_grad():
        w = O.tp_forward("A3GC", x, sds)[2]
    assert float((y - w).norm() / w.norm()) < 2e-6
    # the adjacency parameters were de-aliased: 18 distinct buffers per net
    ptrs = {p.data_ptr() for n, p in nets[0].named_parameters() if "adj" in n}
    assert len(ptrs) == 18


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference`: rank 0 prints one JSON line with impl / cpu_baseline / e2e; other ranks print nothing."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
    sys.path.insert(0, ROOT)
    import bench
    rec = bench.cpu_baseline_record(bench.cpu_reference("A3GC", (1,), timed=1, warmup=0))
    assert rec["kind"] in ("reference", "port") and rec["cores"] >= 1 and rec["value"] > 0 and "B=1" in rec["points"]
    json.dumps(rec)
