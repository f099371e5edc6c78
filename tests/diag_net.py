"""Diagnostic: whole-net TC vs SIMT for a few shapes and image hand-off modes (A3GC_TC_IMG bits)."""
import os, subprocess, sys
CHILD = r'''
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import net_oracle as O
from util import build_net
variant, f0, out, H, B, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
nira = torch.load("tests/golden/nira_template_15_norm.pt")
sd = O.random_state_dict(variant, f0, out, H, nira, seed=3)
x = torch.randn(B, T, 15, f0, generator=torch.Generator().manual_seed(1)).cuda()
y0, h0 = build_net(variant, f0, out, H, sd, nira, engine="simt")(x)
y1, h1 = build_net(variant, f0, out, H, sd, nira, engine="tc")(x)
torch.cuda.synchronize()
rel = lambda a, b: float((a - b).norm() / b.norm())
print(f"{variant} f0={f0} out={out} H={H} B={B} T={T} IMG={__import__('os').environ.get('A3GC_TC_IMG','3')}: y {rel(y1, y0):.2e}  hT {rel(h1[0][0], h0[0][0]):.2e}/{rel(h1[1][0], h0[1][0]):.2e}")
'''
for cfg in (("AAGC", 15, 3, 64, 2, 6), ("A3GC", 15, 3, 64, 2, 6), ("AAGC", 15, 9, 128, 2, 6), ("A3GC", 12, 3, 256, 2, 4)):
    for mode in ("0", "1", "2", "3"):
        env = dict(os.environ, A3GC_TC_IMG=mode)
        r = subprocess.run([sys.executable, "-c", CHILD] + [str(c) for c in cfg], capture_output=True, text=True, timeout=120, env=env)
        print(r.stdout.strip() or r.stderr.strip().splitlines()[-1], flush=True)
