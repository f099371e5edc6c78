"""Diagnostic (not a pytest file): tensor-core engine vs SIMT engine on the same GPU, one child process
per configuration with a timeout so that a deadlocked kernel cannot hang the box.
    python tests/diag_tc.py [filter]
"""
import subprocess
import sys

CHILD = r'''
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import a3gc_ip_b200 as A
variant, H, F, B, T, prec = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
nira = torch.load("tests/golden/nira_template_15_norm.pt").float()
cls = {"AAGC": A.BiAAGC_LSTM, "A3GC": A.BiA3GC_LSTM, "AGC": A.BiAGC_LSTM}[variant]
g = torch.Generator().manual_seed(5)
layer = cls(F, H, nira, activation_fn="tanh")
for n_, p_ in layer.named_parameters():
    if p_.shape == (15, 15): p_.data += 0.05 * torch.randn(15, 15, generator=g)
    elif p_.dim() == 1: p_.data = 0.1 * torch.randn(p_.shape, generator=g)
layer = layer.cuda().eval()
x = torch.randn(B, T, 15, F, generator=g).cuda()
mk = lambda: (0.3 * torch.randn(B, 15, H, generator=g)).cuda()
st = [(mk(), mk()), (mk(), mk())]
layer.set_engine("simt")
y0, s0 = layer(x, st)
layer.set_engine("tc", prec)
y1, s1 = layer(x, st)
torch.cuda.synchronize()
rel = lambda a, b: float((a - b).norm() / b.norm())
H_ = H
print(f"{variant} H={H} F={F} B={B} T={T} {prec}: y fwd {rel(y1[..., :H_], y0[..., :H_]):.2e} rev {rel(y1[..., H_:], y0[..., H_:]):.2e} "
      f"h {rel(s1[0][0], s0[0][0]):.2e}/{rel(s1[1][0], s0[1][0]):.2e} c {rel(s1[0][1], s0[0][1]):.2e}/{rel(s1[1][1], s0[1][1]):.2e} "
      f"finite={bool(torch.isfinite(y1).all())}")
'''

CONFIGS = [
    ("AAGC", 64, 64, 8, 1, "fp32"), ("AAGC", 64, 64, 3, 4, "fp32"), ("A3GC", 64, 64, 8, 1, "fp32"), ("A3GC", 64, 128, 11, 6, "fp32"),
    ("AGC", 64, 64, 8, 3, "fp32"), ("AAGC", 128, 128, 8, 2, "fp32"), ("A3GC", 128, 256, 9, 5, "fp32"),
    ("A3GC", 256, 256, 8, 2, "fp32"), ("A3GC", 256, 512, 17, 6, "fp32"), ("A3GC", 64, 128, 11, 6, "bf16"), ("A3GC", 256, 512, 17, 6, "bf16"),
]

if __name__ == "__main__":
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    for cfg in CONFIGS:
        if flt and flt not in " ".join(map(str, cfg)):
            continue
        try:
            r = subprocess.run([sys.executable, "-c", CHILD] + [str(c) for c in cfg], capture_output=True, text=True, timeout=90)
            tail = r.stderr.strip().splitlines()[-1] if r.returncode and r.stderr.strip() else "ok"
            print((r.stdout.strip() or f"{cfg}: (no stdout)"), "|", tail, flush=True)
        except subprocess.TimeoutExpired:
            print(f"{cfg}: TIMEOUT (deadlock?)", flush=True)
