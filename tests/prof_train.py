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
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    y, _ = net(x); loss = crit(y.view(tgt.shape), tgt); opt.zero_grad(); loss.backward(); opt.step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=int(sys.argv[6]) if len(sys.argv) > 6 else 14, max_name_column_width=60))
