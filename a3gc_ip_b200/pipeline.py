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

    def _chain(self, x: Tensor, slot: int) -> Tuple[Tensor, Tensor, Tensor]:
        y1, _ = self.net1(x, None, slot)
        y2, _ = self.net2(concat_stage_input(x, y1), None, slot)
        y3, _ = self.net3(concat_stage_input(x, y2), None, slot)
        return y1, y2, y3

    @torch.no_grad()
    def forward(self, x: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        """Sequences are independent, so with ``streams`` > 1 the batch is cut into that many contiguous chunks whose
        three-stage chains run concurrently on separate CUDA streams: a chunk's small stages (cluster size 1 / 2) fill
        the SMs that the 4-CTA clusters of another chunk's H = 256 stage cannot use, and partial last waves overlap."""
        per = self._macro_batch(x.shape[0], x.shape[1])
        if x.shape[0] > per:
            parts = [self.forward(x[i:i + per]) for i in range(0, x.shape[0], per)]
            return tuple(torch.cat([o[k] for o in parts], dim=0) for k in range(3))
        n = min(self.streams, max(1, x.shape[0] // 8))
        if n <= 1 or not x.is_cuda:
            return self._chain(x, 0)
        dev = x.device
        main = torch.cuda.current_stream(dev)
        side = self._side.setdefault(dev, [])
        while len(side) < n:
            side.append(torch.cuda.Stream(device=dev))
        # chunk boundaries on multiples of the 8-sequence batch tile
        tiles = (x.shape[0] + 7) // 8
        bounds = [min(x.shape[0], 8 * ((tiles * i) // n)) for i in range(n + 1)]
        ready = torch.cuda.Event()
        ready.record(main)
        outs = []
        for i in range(n):
            st = side[i]
            st.wait_event(ready)
            with torch.cuda.stream(st):
                xi = x[bounds[i]:bounds[i + 1]]
                outs.append(self._chain(xi, i))
        for i in range(n):
            main.wait_stream(side[i])
            for t in outs[i]:
                t.record_stream(main)
        return tuple(torch.cat([o[k] for o in outs], dim=0) for k in range(3))

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

    @torch.no_grad()
    def forward_raw(self, ori: Tensor, acc: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        return self.forward(prepare_input(ori, acc, self.stats))

    @torch.no_grad()
    def forward_host(self, x_host: Tensor, out_host: Optional[Tensor] = None, device: Optional[torch.device] = None) -> Tensor:
        """End-to-end call with HOST buffers: H2D copy of x, three stages, D2H copy of the pose.  With ``streams`` > 1 every
        batch chunk does its own H2D -> chain -> D2H on its stream, so the copies of one chunk overlap the kernels of another
        (pinned host buffers needed for the overlap; pageable ones still work)."""
        device = device or next(self.parameters()).device
        B = x_host.shape[0]
        if out_host is None:
            out_host = torch.empty(B, x_host.shape[1], 15, 9, dtype=torch.float32, pin_memory=True)
        per = self._macro_batch(B, x_host.shape[1])
        if B > per:
            for i in range(0, B, per):
                self.forward_host(x_host[i:i + per], out_host[i:i + per], device)
            return out_host
        n = min(self.streams, max(1, B // 8))
        if n <= 1:
            x = x_host.to(device, non_blocking=True)
            _, _, y3 = self._chain(x, 0)
            out_host.copy_(y3, non_blocking=True)
            torch.cuda.current_stream(device).synchronize()
            return out_host
        main = torch.cuda.current_stream(device)
        side = self._side.setdefault(device, [])
        while len(side) < n:
            side.append(torch.cuda.Stream(device=device))
        tiles = (B + 7) // 8
        bounds = [min(B, 8 * ((tiles * i) // n)) for i in range(n + 1)]
        ready = torch.cuda.Event()
        ready.record(main)
        keep = []
        for i in range(n):
            st = side[i]
            st.wait_event(ready)
            with torch.cuda.stream(st):
                xi = x_host[bounds[i]:bounds[i + 1]].to(device, non_blocking=True)
                y3 = self._chain(xi, i)[2]
                out_host[bounds[i]:bounds[i + 1]].copy_(y3, non_blocking=True)
                keep.append((xi, y3))
        for i in range(n):
            side[i].synchronize()
        return out_host
