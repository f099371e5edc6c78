"""Three-stage "TP" inference chain of evaluate_a3gc_tp.py:64-94, 164-172 on the device.

    x  = prepare_input(ori, acc)                 # normalise, drop IMU 6, scatter to nodes [3,4,13,14,10]
    y1 = net1(x)                                 # leaf joint positions   [B,T,15,3]
    y2 = net2(cat(x, y1))                        # full joint positions   [B,T,15,3]
    y3 = net3(cat(x, y2))                        # reduced global pose    [B,T,15,9]

The reference runs this at B=1 per recording; here B is arbitrary (independent sequences).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib

INPUT_JOINTS = [3, 4, 13, 14, 10]          # evaluate_a3gc_tp.py:65
NUM_NODES = 15


def prepare_input(ori: Tensor, acc: Tensor, stats: Optional[dict] = None, out: Optional[Tensor] = None) -> Tensor:
    """Device version of ``prepare_input`` (evaluate_a3gc_tp.py:64-94).

    ori [..., 54], acc [..., 18] (CUDA fp32) -> [..., 15, 12].  ``stats`` is the dict of
    ``data/all*_train_stats.pt`` (``--norm``); None skips the normalisation.
    """
    ori = _lib.require_cuda_f32(ori, "ori")
    acc = _lib.require_cuda_f32(acc, "acc")
    if ori.shape[-1] != 54 or acc.shape[-1] != 18 or ori.shape[:-1] != acc.shape[:-1]:
        raise RuntimeError(f"prepare_input expects ori [...,54] and acc [...,18], got {tuple(ori.shape)} / {tuple(acc.shape)}")
    lead = tuple(ori.shape[:-1])
    frames = ori.numel() // 54
    if out is None:
        out = torch.empty(*lead, NUM_NODES, 12, dtype=torch.float32, device=ori.device)
    dev = ori.device
    if stats is not None:
        am = stats["acc"]["mean_channel"].to(dev, torch.float32).contiguous()
        asd = stats["acc"]["std_channel"].to(dev, torch.float32).contiguous()
        om = stats["ori"]["mean_channel"].to(dev, torch.float32).contiguous()
        osd = stats["ori"]["std_channel"].to(dev, torch.float32).contiguous()
        ptrs = (am.data_ptr(), asd.data_ptr(), om.data_ptr(), osd.data_ptr())
    else:
        ptrs = (None, None, None, None)
    with torch.cuda.device(dev):
        rc = _lib.lib().a3gc_prepare_input(acc.data_ptr(), ori.data_ptr(), *ptrs, out.data_ptr(), frames, 12, _lib.stream_ptr(dev))
    _lib.check(rc, "a3gc_prepare_input")
    return out


def concat_stage_input(x: Tensor, pos: Tensor) -> Tensor:
    """``torch.cat((x, pos.view(B, T, 15, 3)), dim=-1)`` (evaluate_a3gc_tp.py:168, 170) as one kernel."""
    x = _lib.require_cuda_f32(x, "x")
    pos = _lib.require_cuda_f32(pos, "pos")
    B, T = x.shape[0], x.shape[1]
    out = torch.empty(B, T, NUM_NODES, 15, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().a3gc_concat_stage_input(x.data_ptr(), pos.data_ptr(), out.data_ptr(), B * T, _lib.stream_ptr(x.device))
    _lib.check(rc, "a3gc_concat_stage_input")
    return out


def reduced_global_to_full_local(pose: Tensor, rotsize: int = 9) -> Tensor:
    """``PoseNet3._reduced_glb_to_full_local_mat`` (net_aagc.py:795-800; ``rotsize=6``: :788-793) on the device:
    reduced global pose [..., 15, rotsize] (or [..., 15, 3, 3]) -> full local pose [N, 24, 3, 3]."""
    p = _lib.require_cuda_f32(pose, "pose")
    if rotsize not in (6, 9):
        raise ValueError("rotsize must be 6 or 9")
    frames = p.numel() // (15 * rotsize)
    if frames * 15 * rotsize != p.numel():
        raise RuntimeError(f"pose of shape {tuple(pose.shape)} does not hold [N, 15, {rotsize}]")
    out = torch.empty(frames, 24, 3, 3, dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        rc = _lib.lib().a3gc_reduced_to_full_local(p.data_ptr(), out.data_ptr(), frames, rotsize, _lib.stream_ptr(p.device))
    _lib.check(rc, "a3gc_reduced_to_full_local")
    return out


class TPPipeline(torch.nn.Module):
    """net1 (12 -> 3), net2 (15 -> 3), net3 (15 -> 9) chained as evaluate_a3gc_tp.py:164-172."""

    def __init__(self, net1: torch.nn.Module, net2: torch.nn.Module, net3: torch.nn.Module, stats: Optional[dict] = None,
                 streams: int = 1, max_frames: int = 400_000):
        """``max_frames`` bounds the frames (sequences x steps) in flight: the per-frame workspace of the three stages is
        ~150 KB (operand images and inter-layer activations), so a larger call is cut into sequential macro-batches of
        whole sequences that reuse the same workspaces (the default keeps BASELINE cfg 2, 307 200 frames, in one pass and
        lets cfg 4, 2.46 M frames, run on one 180 GB GPU)."""
        super().__init__()
        self.net1, self.net2, self.net3 = net1, net2, net3
        self.stats = stats
        self.streams = max(1, int(streams))
        self.max_frames = int(max_frames)
        self._side = {}

    def _macro_batch(self, B: int, T: int) -> int:
        """Sequences per macro-batch: a multiple of the 8-sequence batch tile, at least one tile."""
        return max(8, (self.max_frames // max(T, 1)) // 8 * 8)

    def _chain(self, x: Tensor, slot: int, outs=None) -> Tuple[Tensor, Tensor, Tensor]:
        o = outs or (None, None, None)
        y1, _ = self.net1(x, None, slot, o[0])
        y2, _ = self.net2(concat_stage_input(x, y1), None, slot, o[1])
        y3, _ = self.net3(concat_stage_input(x, y2), None, slot, o[2])
        return y1, y2, y3

    def _chain_raw(self, acc: Tensor, ori: Tensor, slot: int, outs=None) -> Tuple[Tensor, Tensor, Tensor]:
        """The same chain from the raw IMU frame: prepare_input and both stage concatenations are fused into the three
        linear_in loads (``a3gc_net_forward_raw``); nothing of shape [B,T,15,12] / [B,T,15,15] is materialised."""
        o = outs or (None, None, None)
        y1, _ = self.net1.forward_raw(acc, ori, self.stats, None, None, slot, o[0])
        y2, _ = self.net2.forward_raw(acc, ori, self.stats, y1, None, slot, o[1])
        y3, _ = self.net3.forward_raw(acc, ori, self.stats, y2, None, slot, o[2])
        return y1, y2, y3

    def _outputs(self, B: int, T: int, dev) -> Tuple[Tensor, Tensor, Tensor]:
        return tuple(torch.empty(B, T, NUM_NODES, n.linear_out.gcn_kernel.shape[0], dtype=torch.float32, device=dev)
                     for n in (self.net1, self.net2, self.net3))

    def _run_chunks(self, B: int, T: int, dev, chain, outs):
        """Run ``chain(lo, hi, slot, out_views)`` over the batch: sequential macro-batches of whole sequences (bounded
        workspace), each cut into ``streams`` contiguous chunks whose three-stage chains run concurrently on separate
        CUDA streams: a chunk's small stages (cluster size 1 / 2) fill the SMs that the 4-CTA clusters of another chunk's
        H = 256 stage cannot use, and partial last waves overlap.  Every chunk writes its slice of the preallocated
        outputs (no concatenation pass)."""
        per = self._macro_batch(B, T)
        for m0 in range(0, B, per):
            m1 = min(B, m0 + per)
            n = min(self.streams, max(1, (m1 - m0) // 8))
            if n <= 1:
                chain(m0, m1, 0, tuple(o[m0:m1] for o in outs))
                continue
            main = torch.cuda.current_stream(dev)
            side = self._side.setdefault(dev, [])
            while len(side) < n:
                side.append(torch.cuda.Stream(device=dev))
            tiles = (m1 - m0 + 7) // 8                               # chunk boundaries on multiples of the 8-sequence batch tile
            bounds = [min(m1, m0 + 8 * ((tiles * i) // n)) for i in range(n + 1)]
            ready = torch.cuda.Event()
            ready.record(main)
            for i in range(n):
                side[i].wait_event(ready)
                with torch.cuda.stream(side[i]):
                    chain(bounds[i], bounds[i + 1], i, tuple(o[bounds[i]:bounds[i + 1]] for o in outs))
            for i in range(n):
                main.wait_stream(side[i])

    @torch.no_grad()
    def forward(self, x: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        """x [B, T, 15, 12] (the output of ``prepare_input``) -> (leaf positions, joint positions, reduced global pose)."""
        x = _lib.require_cuda_f32(x, "x")
        B, T, dev = x.shape[0], x.shape[1], x.device
        outs = self._outputs(B, T, dev)
        self._run_chunks(B, T, dev, lambda lo, hi, slot, o: self._chain(x[lo:hi], slot, o), outs)
        return outs

    @torch.no_grad()
    def forward_raw(self, ori: Tensor, acc: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        """ori [B, T, 54], acc [B, T, 18] (raw IMU frames, argument order of the reference's ``prepare_input(oris, accs)``)
        -> the three stage outputs; normalisation (``self.stats``), IMU drop, node scatter and the stage concatenations
        happen inside the first kernel of every stage."""
        ori = _lib.require_cuda_f32(ori, "ori")
        acc = _lib.require_cuda_f32(acc, "acc")
        B, T, dev = acc.shape[0], acc.shape[1], acc.device
        outs = self._outputs(B, T, dev)
        self._run_chunks(B, T, dev, lambda lo, hi, slot, o: self._chain_raw(acc[lo:hi], ori[lo:hi], slot, o), outs)
        return outs

    @staticmethod
    def window_plan(total: int, window: int, hop: int):
        """Sliding windows over a recording of ``total`` frames: window starts (the last one is shifted back so that it ends
        at the final frame) and, per window, the half-open range [lo, hi) of recording frames whose output is taken from it
        -- the ``hop`` frames around the window centre, the first / last window also covering the head / tail."""
        if window <= 0 or hop <= 0 or hop > window:
            raise ValueError(f"window_plan: need 0 < hop <= window, got window={window} hop={hop}")
        if total <= window:
            return [0], [(0, total)]
        starts = list(range(0, total - window, hop)) + [total - window]
        keep = []
        for i, s0 in enumerate(starts):
            lo = 0 if i == 0 else keep[-1][1]
            hi = total if i == len(starts) - 1 else min(total, s0 + (window + hop) // 2)
            keep.append((lo, max(lo, hi)))
        return starts, keep

    @torch.no_grad()
    def forward_windowed(self, x: Tensor, window: int, hop: int) -> Tensor:
        """Windowed (near-online) inference of ONE long recording x [T, 15, 12] -> pose [T, 15, 9].

        The nets are bidirectional, so the reference's online plumbing (PoseNet3.reset / rnn_state, net_aagc.py:802-812) cannot
        stream; the product form is a sliding window: every window of ``window`` frames is an independent sequence, all
        windows run as ONE batch through the three stages, and each output frame is taken from the window in which it is
        most central (``window_plan``).  ``window`` >= T reproduces the offline result exactly; the latency of a live
        deployment is window/2 + hop frames plus one chain evaluation (31 ms at window = 300 on one B200)."""
        if x.dim() != 3 or x.shape[1] != NUM_NODES:
            raise RuntimeError(f"forward_windowed expects one recording [T, 15, F], got {tuple(x.shape)}")
        x = _lib.require_cuda_f32(x, "x")
        total = x.shape[0]
        starts, keep = self.window_plan(total, window, hop)
        w = min(window, total)
        idx = torch.tensor(starts, device=x.device).unsqueeze(1) + torch.arange(w, device=x.device).unsqueeze(0)
        y3 = self.forward(x[idx])[2]                                   # [windows, w, 15, 9]
        out = torch.empty(total, NUM_NODES, y3.shape[-1], dtype=torch.float32, device=x.device)
        for i, (s0, (lo, hi)) in enumerate(zip(starts, keep)):
            if hi > lo:
                out[lo:hi] = y3[i, lo - s0:hi - s0]
        return out

    def _host_chunks(self, B: int, T: int, device, chain):
        """Host-buffer driver: every batch chunk does its own H2D -> chain -> D2H on its stream, so the copies of one chunk
        overlap the kernels of another (pinned host buffers needed for the overlap; pageable ones still work)."""
        per = self._macro_batch(B, T)
        for m0 in range(0, B, per):
            m1 = min(B, m0 + per)
            n = min(self.streams, max(1, (m1 - m0) // 8))
            main = torch.cuda.current_stream(device)
            if n <= 1:
                keep = chain(m0, m1, 0)
                main.synchronize()
                continue
            side = self._side.setdefault(device, [])
            while len(side) < n:
                side.append(torch.cuda.Stream(device=device))
            tiles = (m1 - m0 + 7) // 8
            bounds = [min(m1, m0 + 8 * ((tiles * i) // n)) for i in range(n + 1)]
            ready = torch.cuda.Event()
            ready.record(main)
            keep = []
            for i in range(n):
                side[i].wait_event(ready)
                with torch.cuda.stream(side[i]):
                    keep.append(chain(bounds[i], bounds[i + 1], i))
            for i in range(n):
                side[i].synchronize()

    @torch.no_grad()
    def forward_host(self, x_host: Tensor, out_host: Optional[Tensor] = None, device: Optional[torch.device] = None) -> Tensor:
        """End-to-end call with HOST buffers holding the prepared input x [B, T, 15, 12] (720 B per frame): H2D copy of x,
        three stages, D2H copy of the pose [B, T, 15, 9]."""
        device = device or next(self.parameters()).device
        B, T = x_host.shape[0], x_host.shape[1]
        if out_host is None:
            out_host = torch.empty(B, T, 15, 9, dtype=torch.float32, pin_memory=True)

        def chain(lo, hi, slot):
            xi = x_host[lo:hi].to(device, non_blocking=True)
            y3 = self._chain(xi, slot)[2]
            out_host[lo:hi].copy_(y3, non_blocking=True)
            return xi, y3

        with torch.cuda.device(device):
            self._host_chunks(B, T, device, chain)
        return out_host

    @torch.no_grad()
    def forward_host_raw(self, ori_host: Tensor, acc_host: Tensor, out_host: Optional[Tensor] = None,
                         device: Optional[torch.device] = None) -> Tensor:
        """End-to-end call with HOST buffers holding the raw IMU frames (ori [B, T, 54], acc [B, T, 18]: 288 B per frame, what
        ``evaluate_a3gc_tp.py`` reads from ``test_tp.pt``): H2D of the raw frames, fused prepare_input + three stages, D2H of
        the pose [B, T, 15, 9]."""
        device = device or next(self.parameters()).device
        B, T = acc_host.shape[0], acc_host.shape[1]
        if out_host is None:
            out_host = torch.empty(B, T, 15, 9, dtype=torch.float32, pin_memory=True)

        def chain(lo, hi, slot):
            ai = acc_host[lo:hi].to(device, non_blocking=True)
            oi = ori_host[lo:hi].to(device, non_blocking=True)
            y3 = self._chain_raw(ai, oi, slot)[2]
            out_host[lo:hi].copy_(y3, non_blocking=True)
            return ai, oi, y3

        with torch.cuda.device(device):
            self._host_chunks(B, T, device, chain)
        return out_host

    def cache_packed_weights(self, on: bool = True):
        """Serving with frozen weights: keep the packed weights of the three nets across calls (see ``_Net.cache_packed_weights``)."""
        for n in (self.net1, self.net2, self.net3):
            n.cache_packed_weights(on)
        return self

    def release_workspaces(self) -> None:
        """Free the scratch memory (operand images, inter-layer activations: ~150 KB per frame in flight) of the three nets."""
        for n in (self.net1, self.net2, self.net3):
            n.release_workspaces()
