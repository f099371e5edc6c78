"""Timing of the hoisted weight-gradient GEMM shapes of one stage-1 training step (dW = dzm^T [x | h_prev], R = B T 15 rows)
in the forms torch / cuBLAS offers: TF32 on fp32 operands (what training.py issues, three passes), bf16 operands, a
pre-transposed left operand, and one GEMM over the concatenated right operand.  python tests/diag_dw_gemm.py"""
import torch

torch.backends.cuda.matmul.fp32_precision = "tf32"
R, M = 256 * 200 * 15, 1024


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    a = torch.randn(R, M, device="cuda")
    at = a.t().contiguous()
    ab = a.bfloat16()
    for N in (256, 512, 768):
        b = torch.randn(R, N, device="cuda")
        bb = b.bfloat16()
        out = torch.empty(M, N, device="cuda")
        gf = 2.0 * R * M * N / 1e9
        rows = [("tf32 a.t() @ b", lambda: torch.mm(a.t(), b, out=out)),
                ("tf32 b.t() @ a (transposed result)", lambda: torch.mm(b.t(), a)),
                ("tf32 at @ b (left operand stored [M, R])", lambda: torch.mm(at, b, out=out)),
                ("bf16 a.t() @ b", lambda: torch.mm(ab.t(), bb)),
                ("transpose copy of a", lambda: a.t().contiguous())]
        for name, fn in rows:
            ms = timed(fn)
            print(f"N={N:4d} {name:45s} {ms:7.3f} ms  {gf / ms:8.1f} TFLOP/s-equivalent" if "copy" not in name else f"N={N:4d} {name:45s} {ms:7.3f} ms", flush=True)


if __name__ == "__main__":
    main()
