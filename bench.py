#!/usr/bin/env python
"""bench.py -- A3GC-TP frames/s on B200 (BASELINE.json metric), roofline fraction and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (three-stage A3GC-TP forward: prepare-free node input
[B,T,15,12] -> leaf positions -> joint positions -> reduced global pose [B,T,15,9]) over one batch
of synthetic IMU sequences.  Workload = BASELINE.json configs[1]: A3GC-TP fp32, B=1024 x T=300 per
GPU (weak scaling: every rank owns its own 1024 independent sequences, no data-path collective).
Weights: stage 1 (H=256) random-init seed 0 (its checkpoint is not shipped), stages 2-3 (H=64, 128)
from the reference's trained_models/A3GC (committed as fixtures under tests/golden/weights).

Reported on ONE JSON line (rank 0):
  value     frames/s with inputs resident in HBM (CUDA-event timed, max over ranks)
  e2e       same metric through the public API with HOST buffers (pinned H2D of x + D2H of the pose per step)
  roofline  tensor-core roofline of the dominant kernel (stage-1 rnn2 layer launch), timed live with CUDA events
  cpu_baseline  the CPU oracle port (oracle/net_oracle.py, same algorithm as the reference) on the host cores
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

METRIC = "A3GC-TP frames/sec"
UNIT = "frames/s"
B_PER_GPU, T_STEPS = 1024, 300
MFLOP_PER_FRAME = 115.70          # algorithmic dense MFLOP per frame, A3GC-TP (SURVEY.md 8d / BASELINE.md 4)
WORKLOAD = "A3GC-TP forward fp32, B=1024 x T=300 per GPU, 15-node graph, stages H=256/64/128 (BASELINE cfg 2)"


def load_nira():
    return torch.load(os.path.join(ROOT, "tests", "golden", "nira_template_15_norm.pt"))


def tp_weights(nira):
    from util import tp_state_dicts
    return tp_state_dicts("A3GC", nira)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_fps(nira, sample_b, steps, warmup):
    """The reference's CPU algorithm (oracle port) on this box's host cores; returns (frames/s, cores, ms/step)."""
    from oracle import net_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sds = tp_weights(nira)
    x = O.synthetic_input(sample_b, T_STEPS, seed=1234)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.tp_forward("A3GC", x, sds)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    tot = sum(times)
    return sample_b * T_STEPS * len(times) / tot, cores, 1e3 * tot / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    nira = load_nira()
    sample_b = 16
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    fps, cores, ms = cpu_reference_fps(nira, sample_b, steps, warmup)
    sample = f"oracle port of net_aagc.py (torch CPU eager), B={sample_b} x T={T_STEPS} sequences per step, {steps} timed steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_train(args):
    """Secondary workload (BASELINE cfg 5, not the headline line): A3GC-TP training step -- for each of the three
    stages one optimisation step as train_a3gc_tp.py:74-84 does it (train-mode forward with the reference's dropout,
    pose loss, BPTT backward, Adam), B=256 x T=200 per GPU, teacher-forced synthetic inputs; data-parallel ranks
    all-reduce one flat fp32 gradient bucket per stage over NCCL.  frames/s = world * B * T / step time."""
    import torch.distributed as dist
    import a3gc_ip_b200 as A
    world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, T = (args.batch if args.batch != B_PER_GPU else 256), 200
    W, K = max(args.warmup, 3), max(args.steps, 1)
    nira = load_nira().float()
    torch.manual_seed(0)                                   # identical initial weights on every rank
    shapes = ((12, 3, 256), (15, 3, 64), (15, 9, 128))
    nets = [A.A3GC_net(f0, o, h, nira).to(dev).train() for f0, o, h in shapes]
    opts = [torch.optim.Adam(n.parameters(), lr=1e-3) for n in nets]
    reds = [A.FlatGradAllReducer(n.parameters()) if world > 1 else None for n in nets]
    crit = A.pose_loss()
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    xs = [torch.randn(B, T, 15, f0, generator=g, device=dev) for f0, _, _ in shapes]
    ts = [torch.randn(B, T, 15 * o, generator=g, device=dev) for _, o, _ in shapes]

    # The three stage models are independent (the reference trains each with its own run of train_a3gc_tp.py), so their
    # optimisation steps are enqueued on three CUDA streams: the small stages fill the SMs that the H=256 stage's partial
    # waves (reverse-time chain: 256 CTAs on 148 SMs; tcgen05 forward: 33 clusters of 4) leave idle.  --train-streams 1
    # runs them one after the other.
    conc = args.train_streams > 1
    side = [torch.cuda.Stream(device=dev) for _ in nets] if conc else None

    def step():
        if not conc:
            return [A.train_step(n, crit, o, x, t, r) for n, o, x, t, r in zip(nets, opts, xs, ts, reds)]
        main = torch.cuda.current_stream(dev)
        ready = torch.cuda.Event()
        ready.record(main)
        out = []
        for st, n, o, x, t, r in zip(side, nets, opts, xs, ts, reds):
            st.wait_event(ready)
            with torch.cuda.stream(st):
                out.append(A.train_step(n, crit, o, x, t, r))
        for st in side:
            main.wait_stream(st)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(W):
        step()
    barrier()
    L = A.lib(); L.a3gc_reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        losses = step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    barrier()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        msk = float(ms.item()) / K
        print(json.dumps({
            "metric": "A3GC-TP train frames/sec", "value": world * B * T / (msk / 1e3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": msk, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "A3GC-TP training step (fwd + BPTT bwd + Adam, 3 stages), B=%d x T=%d per GPU (BASELINE cfg 5)" % (B, T),
                       "dropout": "reference defaults 0.2 / 0.3 / 0.3", "stage_streams": 3 if conc else 1, "allreduce_bytes_per_step": sum(r.nbytes for r in reds if r) or 0,
                       "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30},
            "gpu_launches": int(L.a3gc_launch_count()), "losses": [float(l) for l in losses],
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default=os.environ.get("A3GC_ENGINE", "auto"))
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="sequences per GPU (default: the BASELINE cfg-2 value)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=int(os.environ.get("A3GC_STREAMS", 4)),
                    help="batch chunks whose three-stage chains run concurrently on separate CUDA streams")
    ap.add_argument("--train-streams", type=int, default=3, help="--workload train: 3 = the three independent stage steps on three CUDA streams, 1 = sequential")
    ap.add_argument("--variant", default="A3GC", choices=["A3GC", "AAGC", "AGC", "GGRU"],
                    help="cell family of the three-stage pipeline (headline: A3GC; the others are the cfg 3 / cfg 4 side lines)")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"], help="fp32 = parity path (headline); bf16 = cfg-3 path")
    ap.add_argument("--seq-len", type=int, default=T_STEPS)
    ap.add_argument("--workload", default="infer", choices=["infer", "train"],
                    help="infer = the headline line (BASELINE cfg 2); train = the secondary cfg-5 training-step line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "train":
        return run_train(args)

    import torch.distributed as dist
    import a3gc_ip_b200 as A
    from oracle import net_oracle as O
    from util import build_tp

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, K = max(args.warmup, 3), max(args.steps, 1)
    B = args.batch
    nira = load_nira()
    T = args.seq_len
    pipe, _ = build_tp(args.variant, nira, device=dev, engine=args.engine, precision=args.precision)
    pipe.streams = args.streams
    mflop = {"A3GC": 115.70, "AGC": 115.70, "AAGC": 103.95, "GGRU": 88.46}[args.variant]
    headline = args.variant == "A3GC" and args.precision == "fp32" and T == T_STEPS and B == B_PER_GPU
    workload = WORKLOAD if headline else f"{args.variant}-TP forward {args.precision}, B={B} x T={T} per GPU (side line, not the headline config)"
    L = A.lib()

    # synthetic inputs: seed 1234 + rank; x is 221 MB (> 126 MB L2), pose output 166 MB
    x_host = O.synthetic_input(B, T, seed=1234 + rank).pin_memory()
    y_host = torch.empty(B, T, 15, 9, dtype=torch.float32).pin_memory()
    x = x_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        barrier()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(W):
        pipe(x)
    torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    L.a3gc_reset_launch_count()
    ms_total = timed(lambda: pipe(x), K)
    launches = int(L.a3gc_launch_count())
    clocks = sampler.stop() if rank == 0 else None

    # end to end through the public API with host buffers (H2D of x and D2H of the pose inside the timed region)
    pipe.forward_host(x_host, y_host, dev)
    ms_e2e = timed(lambda: pipe.forward_host(x_host, y_host, dev), K)

    # dominant kernel: per-launch CUDA-event times of the recurrent layer launches
    L.a3gc_profile_enable(1)
    pipe.streams = 1                      # each launch timed alone (no co-running chunk on another stream)
    pipe(x)
    pipe.streams = args.streams
    torch.cuda.synchronize(dev)
    recs = []
    for i in range(L.a3gc_profile_count()):
        lab = C.create_string_buffer(96)
        ms, fl = C.c_float(), C.c_double()
        L.a3gc_profile_get(i, lab, 96, C.byref(ms), C.byref(fl))
        recs.append({"kernel": lab.value.decode(), "ms": ms.value, "gflop": fl.value / 1e9})
    L.a3gc_profile_enable(0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    frames_per_step = world * B * T
    fps = frames_per_step * K / (ms_total / 1e3)
    fps_e2e = frames_per_step * K / (ms_e2e / 1e3)
    peaks, src = measured_peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    dom = max(recs, key=lambda r: r["ms"]) if recs else None
    layer_ms = sum(r["ms"] for r in recs)
    roof = None
    if dom:
        ach = dom["gflop"] / dom["ms"]            # GFLOP / ms == TFLOP/s
        tc = dom["kernel"].startswith("tc")
        traffic, tnote = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")     # dram bytes per launch from the committed ncu --set full capture
        if os.path.exists(tp) and headline:
            rec = json.load(open(tp)).get(dom["kernel"])
            if rec:
                traffic, tnote = rec["dram_bytes"], rec["source"]
        roof = {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic, "traffic_source": tnote,
                "kernel": dom["kernel"], "kernel_ms": dom["ms"], "peak_source": f"bf16_tflops_sustained ({src})",
                "note": (("algorithmic FLOPs (2MNK of gate+attention GEMMs); the fp32-parity tensor path executes 3 fp16-split passes, "
                          "so executed tensor FLOPs are 3x" if args.precision == "fp32" else "algorithmic FLOPs; bf16 operands, one tensor pass") if tc else "SIMT fp32 engine: FFMA pipe, quoted against the tensor peak"),
                "layer_share_of_step": layer_ms / (ms_total / K), "launches": recs,
                "whole_step_tflops": mflop * 1e6 * (B * T) / (ms_total / K / 1e3) / 1e12}
    cpu = None
    if not args.no_cpu_baseline and world == 1:       # the CPU baseline is reported at N = 1 only
        sb = 16
        cfps, cores, _ = cpu_reference_fps(nira, sb, 2, 1)
        cpu = {"value": cfps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"oracle port of net_aagc.py (torch CPU eager), B={sb} x T={T_STEPS}, 1 warm-up + 2 timed passes"}
    out = {
        "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
        "config": {"workload": workload, "batch_per_gpu": B, "seq_len": T, "engine": args.engine, "streams": args.streams,
                   "l2": "inputs larger than L2 (x 221 MB, activations GBs per step); no explicit flush",
                   "weights": "stage1 random-init seed 0; stages 2-3 " + (f"trained_models/{'A3GC' if args.variant == 'A3GC' else 'G-GRU'}" if args.variant in ("A3GC", "GGRU") else "random-init (no checkpoints shipped)")},
        "e2e": {"value": fps_e2e, "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4 * world, "d2h_bytes_per_step": y_host.numel() * 4 * world,
                "ms_per_step": ms_e2e / K},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
