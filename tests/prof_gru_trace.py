"""Profiling driver (not a pytest file): per-phase timeline of CTA (0,0) of one bidirectional G-GRU layer on the tensor-core
engine (A3GC_TC_TRACE=gru).    python tests/prof_gru_trace.py H F [B] [T]
"""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
os.environ["A3GC_TC_TRACE"] = "gru"
import a3gc_ip_b200 as A

H, F = int(sys.argv[1]), int(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
T = int(sys.argv[4]) if len(sys.argv) > 4 else 40
nira = torch.load("tests/golden/nira_template_15_norm.pt").float()
layer = A.BiG_GRU(F, H, nira).cuda().eval().set_engine("tc", "fp32")
x = torch.randn(B, T, 15, F, generator=torch.Generator().manual_seed(1)).cuda()
z = torch.zeros(B, 15, H).cuda()
for _ in range(3):
    layer(x, [z, z.clone()])
torch.cuda.synchronize()
buf = np.zeros((2, 16, 16), dtype=np.uint64)
A.lib().a3gc_debug_read_tc_trace(buf.ctypes.data)
ne = ["start", "acc_full", "h_free", "gates_done", "mix_done", "published", "emitted"]
nm = ["start", "h_ready", "hpart_issued", "acc_empty", "xpart_issued"]
print(f"G-GRU H={H} F={F} B={B} T={T}: cycles relative to the epilogue's step start")
for t in range(3, 7):
    e0 = int(buf[0, t, 0])
    print(f"  step {t}: epi " + " ".join(f"{n}={int(buf[0, t, i]) - e0}" for i, n in enumerate(ne)))
    print(f"          mma " + " ".join(f"{n}={int(buf[1, t, i]) - e0}" for i, n in enumerate(nm)))
print(f"  steps 2..12: {(int(buf[0, 12, 0]) - int(buf[0, 2, 0])) // 10} cycles per step")
m = [int(v) for v in buf[1, 15, :3]]
print(f"  MMA thread over the whole launch: {m[2]} cycles, of which waiting for a full ring slot {m[0]} ({100 * m[0] / max(m[2], 1):.0f} %), "
      f"for the state exchange {m[1]} ({100 * m[1] / max(m[2], 1):.0f} %)")
for i in range(3):
    w, tot = int(buf[0, 15, 2 * i]), int(buf[0, 15, 2 * i + 1])
    if tot:
        print(f"  producer {i}: {tot} cycles, of which waiting for an empty slot {w} ({100 * w / tot:.0f} %)")
