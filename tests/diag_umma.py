"""Diagnostic (not a pytest file): run the UMMA self-test variants, each in its own process, and print
the error of each, so that one GPU call tells which descriptor convention the hardware uses.
    python tests/diag_umma.py            # driver
"""
import subprocess
import sys

CHILD = r'''
import sys, torch
sys.path.insert(0, ".")
from a3gc_ip_b200 import _lib
K, N, flags = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dt = torch.bfloat16 if flags & 1 else torch.float16
g = torch.Generator().manual_seed(1)
a = torch.randn(128, K, generator=g).to(dt); b = torch.randn(N, K, generator=g).to(dt)
img = lambda m: m.view(m.shape[0], K // 8, 8).permute(1, 0, 2).contiguous().cuda()
ai, bi = img(a), img(b)
d = torch.full((128, N), float("nan"), device="cuda")
rc = _lib.lib().a3gc_tc_selftest(ai.data_ptr(), bi.data_ptr(), d.data_ptr(), K, N, flags, _lib.stream_ptr(d.device))
torch.cuda.synchronize()
want = a.float() @ b.float().t()
got = d.cpu()
err = float((got - want).abs().max() / want.abs().max())
print(f"K={K} N={N} flags={flags} rc={rc} err={err:.3e} nan={int(torch.isnan(got).sum())} got[0,:4]={got[0,:4].tolist()} want[0,:4]={want[0,:4].tolist()}")
'''

if __name__ == "__main__":
    for K, N, flags in ((16, 64, 0), (16, 64, 2), (64, 256, 0), (64, 256, 2), (128, 128, 1)):
        r = subprocess.run([sys.executable, "-c", CHILD, str(K), str(N), str(flags)], capture_output=True, text=True, timeout=120)
        print((r.stdout.strip() or "(no stdout)"), "|", r.stderr.strip().splitlines()[-1] if r.returncode else "ok")
