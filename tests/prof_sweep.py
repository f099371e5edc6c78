"""In-process sweep driver (not a pytest file): times one bidirectional A3GC layer on the tensor-core engine for a list of
shapes x environment-knob settings (the library reads its A3GC_TC_* knobs with getenv at every launch).
    python tests/prof_sweep.py "H,F;H,F;..." "KNOB=v KNOB=v|KNOB=v|..." [B] [T] [precision] [variant]
"""
import os
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import a3gc_ip_b200 as A

shapes = [tuple(int(v) for v in s.split(",")) for s in sys.argv[1].split(";")]
settings = sys.argv[2].split("|")
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
T = int(sys.argv[4]) if len(sys.argv) > 4 else 40
prec = sys.argv[5] if len(sys.argv) > 5 else "fp32"
variant = sys.argv[6] if len(sys.argv) > 6 else "A3GC"
nira = torch.load("tests/golden/nira_template_15_norm.pt").float()
cls = {"A3GC": A.BiA3GC_LSTM, "AAGC": A.BiAAGC_LSTM, "AGC": A.BiAGC_LSTM, "GGRU": A.BiG_GRU}[variant]
KNOBS = ("A3GC_TC_HSEP", "A3GC_TC_SPLIT", "A3GC_TC_STAGES", "A3GC_TC_EARLYPUB", "A3GC_TC_PUBORDER", "A3GC_TC_TRACE", "A3GC_TC_QVEC",
         "A3GC_TC_XCHG", "A3GC_TC_OPT", "A3GC_TC_NPROD", "A3GC_TC_ACOLL", "A3GC_TC_WSTAGES", "A3GC_TC_XSTAGES", "A3GC_TC_XPREFETCH", "A3GC_TC_XSPLIT", "A3GC_TC_XDEFER")
names_e = ["start", "acc_full", "ep1_done", "h_free", "pub_hhat", "att_full", "q_sent", "att2_full", "ep3_done", "a_ready", "out_done", "pub_h"]
names_m = ["start", "h_ready", "hpart_issued", "xpart_issued", "a1_go", "a1_issued", "a2_go", "a2_issued", "x3_issued"]
for (H, F) in shapes:
    layer = cls(F, H, nira, activation_fn="tanh").cuda().eval().set_engine("tc", prec)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, T, 15, F, generator=g).cuda()
    z = torch.zeros(B, 15, H).cuda()
    st = [z, z.clone()] if variant == "GGRU" else [(z, z.clone()), (z.clone(), z.clone())]
    ref = None
    for setting in settings:
        for k in KNOBS:
            os.environ.pop(k, None)
        for kv in setting.split():
            k, v = kv.split("=")
            os.environ[k] = v
        for _ in range(2):
            layer(x, st)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            y, _ = layer(x, st)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        if ref is None:
            ref = y.clone()
        dev = float((y - ref).norm() / ref.norm())
        fl = (2.0 * 15 * (F + H) * 4 * H + (34.0 * H * H + 30 * H if variant != "AAGC" else 0)) * B * T * 2
        if variant == "GGRU":
            fl = 2.0 * 15 * (3 * F * H + 4 * H * H) * B * T * 2
        print(f"H={H} F={F} B={B} T={T} {variant}/{prec} [{setting}]: {best:.3f} ms  {best * 1e3 / T:.1f} us/step  {fl / best / 1e9:.1f} TFLOP/s  "
              f"finite={bool(torch.isfinite(y).all())} rel-dev-vs-first={dev:.2e}", flush=True)
        if os.environ.get("A3GC_TC_TRACE"):
            import numpy as np
            buf = np.zeros((2, 16, 16), dtype=np.uint64)
            A.lib().a3gc_debug_read_tc_trace(buf.ctypes.data)
            t0 = int(buf[0, 0, 0])
            for t in (3, 4):
                e = [int(v) - t0 for v in buf[0, t, :12]]
                mm = [int(v) - t0 for v in buf[1, t, :9]]
                print(f"   step {t}: epi " + " ".join(f"{n}={v - e[0]}" for n, v in zip(names_e, e)))
                print(f"           mma " + " ".join(f"{n}={v - e[0]}" for n, v in zip(names_m, mm)))
            cyc = int(buf[0, 12, 1]) - int(buf[0, 2, 1])
            print(f"   steps 2..12: {cyc / 10:.0f} cycles per step", flush=True)
            w = [int(v) for v in buf[1, 15, :4]]
            if variant != "GGRU" and w[3]:
                print(f"   MMA warp over the launch: {w[3]} cycles; waiting for a full weight slot {100 * w[0] / w[3]:.0f} %, x slot {100 * w[1] / w[3]:.0f} %, "
                      f"state blocks / q / accumulators {100 * w[2] / w[3]:.0f} %, issuing {100 * (w[3] - w[0] - w[1] - w[2]) / w[3]:.0f} %", flush=True)
