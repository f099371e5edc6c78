"""Profiling driver (not a pytest file): phases of one A3GC training step.  python tests/prof_train.py H F0 OUT B T"""
import sys, time
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import a3gc_ip_b200 as A

H, F0, OUT, B, T = (int(v) for v in sys.argv[1:6])
nira = torch.load("tests/golden/nira_template_15_norm.pt").float()
net = A.A3GC_net(F0, OUT, H, nira).cuda().train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3)
crit = A.pose_loss()
x = torch.randn(B, T, 15, F0, device="cuda"); tgt = torch.randn(B, T, 15 * OUT, device="cuda")
def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(3):
    e0 = ev(); y, _ = net(x); e1 = ev(); loss = crit(y.view(tgt.shape), tgt); opt.zero_grad(); e2 = ev(); loss.backward(); e3 = ev(); opt.step(); e4 = ev()
    torch.cuda.synchronize()
    print(f"iter {it}: fwd {e0.elapsed_time(e1):.1f} ms  bwd {e2.elapsed_time(e3):.1f} ms  adam {e3.elapsed_time(e4):.1f} ms  total {e0.elapsed_time(e4):.1f} ms  mem {torch.cuda.max_memory_allocated()/2**30:.1f} GB")
import os
if os.environ.get("A3GC_TC_TRACE"):
    # phase stamps of the tcgen05 forward in TRAIN mode (the last layer launch of the forward above): cycles within a step
    import numpy as np
    names_e = ["start", "acc_full", "ep1_done", "h_free", "pub_hhat", "att_full", "q_sent", "att2_full", "ep3_done", "a_ready", "out_done", "pub_h"]
    names_m = ["start", "h_ready", "hpart_issued", "xpart_issued", "a1_go", "a1_issued", "a2_go", "a2_issued", "x3_issued"]
    y, _ = net(x); torch.cuda.synchronize()
    buf = np.zeros((2, 16, 16), dtype=np.uint64)
    A.lib().a3gc_debug_read_tc_trace(buf.ctypes.data)
    for t in (3, 4):
        e = [int(v) for v in buf[0, t, :12]]
        mm = [int(v) for v in buf[1, t, :9]]
        print(f"   step {t}: epi " + " ".join(f"{n}={v - e[0]}" for n, v in zip(names_e, e)))
        print(f"           mma " + " ".join(f"{n}={v - e[0]}" for n, v in zip(names_m, mm)))
    print(f"   steps 2..12: {(int(buf[0, 12, 1]) - int(buf[0, 2, 1])) / 10:.0f} cycles per step", flush=True)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    y, _ = net(x); loss = crit(y.view(tgt.shape), tgt); opt.zero_grad(); loss.backward(); opt.step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=int(sys.argv[6]) if len(sys.argv) > 6 else 14, max_name_column_width=60))
