"""Profiling driver (not a pytest file): one bidirectional A3GC layer on the tensor-core engine.
    python tests/prof_tc.py H F B T [engine] [precision]
"""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import a3gc_ip_b200 as A

H, F, B, T = (int(v) for v in sys.argv[1:5])
engine = sys.argv[5] if len(sys.argv) > 5 else "tc"
prec = sys.argv[6] if len(sys.argv) > 6 else "fp32"
nira = torch.load("tests/golden/nira_template_15_norm.pt").float()
layer = A.BiA3GC_LSTM(F, H, nira, activation_fn="tanh").cuda().eval().set_engine(engine, prec)
g = torch.Generator().manual_seed(1)
x = torch.randn(B, T, 15, F, generator=g).cuda()
z = torch.zeros(B, 15, H).cuda()
st = [(z, z.clone()), (z.clone(), z.clone())]
for _ in range(2):
    layer(x, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
y, _ = layer(x, st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
fl = (2.0 * 15 * (F + H) * 4 * H + 34.0 * H * H + 30 * H) * B * T * 2
print(f"H={H} F={F} B={B} T={T} {engine}/{prec}: {ms:.3f} ms  {ms * 1e3 / T:.1f} us/step  {fl / ms / 1e9:.2f} TFLOP/s algorithmic  finite={bool(torch.isfinite(y).all())}")

import os
if os.environ.get("A3GC_TC_TRACE"):
    import ctypes as C, numpy as np
    buf = np.zeros((2, 16, 16), dtype=np.uint64)
    A.lib().a3gc_debug_read_tc_trace(buf.ctypes.data)
    t0 = int(buf[0, 0, 0])
    names_e = ["start", "acc_full", "ep1_done", "h_free", "pub_hhat", "att_full", "q_sent", "att2_full", "ep3_done", "a_ready", "out_done", "pub_h"]
    names_m = ["start", "h_ready", "hpart_issued", "xpart_issued", "a1_go", "a1_issued", "a2_go", "a2_issued"]
    for t in range(2, 8):
        e = [int(v) - t0 for v in buf[0, t, :12]]
        mm = [int(v) - t0 for v in buf[1, t, :8]]
        print(f"step {t}: epi " + " ".join(f"{n}={v - e[0]}" for n, v in zip(names_e, e)) + f" | abs start {e[0]}")
        print(f"        mma " + " ".join(f"{n}={v - e[0]}" for n, v in zip(names_m, mm)))
        x = [int(v) - t0 - e[0] for v in buf[0, t, 12:15]]
        y = [int(v) - t0 - e[0] for v in buf[1, t, 8:13]]
        print(f"        attention GEMM: hy block of peer +1 / +2 / +3 available at {y[1]} / {y[2]} / {y[3]}")
        print(f"        q-step: stores_done={x[0]} fence_done={x[1]} bar_done={x[2]} | ep1 q=1: ld_done={y[0]} sts_done={y[1]} bar1={y[2]} compute_done={y[3]} bar2={y[4]}")
    cyc = int(buf[0, 12, 1]) - int(buf[0, 2, 1]); ns = int(buf[0, 12, 15]) - int(buf[0, 2, 15])
    if ns > 0:
        print(f"steps 2..12 of CTA (0,0): {cyc} cycles in {ns} ns -> SM clock {cyc / ns * 1e3:.0f} MHz while the kernel runs")
