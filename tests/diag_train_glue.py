"""Diagnostic (not a pytest file): which torch expressions of the hoisted-gradient glue launch copy kernels, at stage-1 sizes."""
import sys
import torch
from torch.profiler import profile, ProfilerActivity

TB, H = 256 * 200, 256
dz = torch.randn(TB, 4, H, 16, device="cuda")
u = torch.randn(TB, 4, H, 16, device="cuda")
dep = torch.randn(TB, H, 16, device="cuda")
hh = torch.randn(TB, H, 16, device="cuda")
dap = torch.randn(TB, 16, device="cuda")


def run(name, fn):
    fn(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(); torch.cuda.synchronize()
    print(f"== {name}")
    for e in prof.key_averages():
        print(f"   {e.key[:70]:70s} x{e.count:3d}  {e.self_device_time_total / 1e3:8.3f} ms")


run("dP per gate bmm", lambda: [torch.bmm(dz[:, i].transpose(1, 2), u[:, i]).sum(0) for i in range(4)])
run("dP batched (r,g) bmm", lambda: torch.bmm(dz.reshape(TB * 4, H, 16).transpose(1, 2), u.reshape(TB * 4, H, 16)).reshape(TB, 4, 16, 16).sum(0))
run("dP matmul 4d", lambda: torch.matmul(dz.transpose(2, 3), u).sum(0))
run("wh einsum", lambda: torch.einsum("rkn,rjn->kj", dep, hh))
run("wh permute+mm", lambda: dep.permute(1, 0, 2).reshape(H, TB * 16) @ hh.permute(0, 2, 1).reshape(TB * 16, H))
run("u einsum", lambda: torch.einsum("rn,rjn->j", dap, hh))
run("bias sum", lambda: dz.sum(dim=(0, 3)))
run("dep sum", lambda: dep.sum(dim=(0, 2)))
