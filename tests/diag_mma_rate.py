"""Tuning aid (not a pytest file): UMMA issue rate vs shared-memory operand layout.  python tests/diag_mma_rate.py"""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from a3gc_ip_b200 import _lib

L = _lib.lib()
out = torch.zeros(1, device="cuda")


def run(name, n, lt, a_lbo, a_sbo, b_lbo, b_sbo, a_k, b_k, nk, iters=256, grid=1):
    rc = L.a3gc_tc_mma_bench(n, lt, a_lbo, a_sbo, b_lbo, b_sbo, a_k, b_k, nk, iters, grid, out.data_ptr(), _lib.stream_ptr(out.device))
    _lib.check(rc, "a3gc_tc_mma_bench")
    torch.cuda.synchronize()
    print(f"{name:44s} N={n:3d} nk={nk:2d} grid={grid:3d}: {out.item():7.1f} cycles/MMA (floor {128 * n // 256})")


for grid in (1, 148):
    for n in (256, 128, 64):
        # engine layout: [K/8][rows][16 B]  (LBO = rows*16, SBO = 128), K-step = 2 chunks
        run("no-swizzle, K chunks rows*16 B apart", n, 0, 128 * 16, 128, n * 16, 128, 2 * 128 * 16, 2 * n * 16, 8, grid=grid)
        # [rows/8][K/8 = 2][8 rows][16 B]: both K chunks of an 8-row group adjacent (LBO = 128, SBO = 256)
        run("no-swizzle, K chunks adjacent (256 B groups)", n, 0, 128, 256, 128, 256, 128 * 32, n * 32, 8, grid=grid)
        # 128-byte swizzle: rows of 64 elements, 8-row atoms of 1024 B; K-step = +32 B inside the row
        run("128B swizzle", n, 2, 16, 1024, 16, 1024, 32, 32, 4, grid=grid)
        run("same A/B every MMA (no-swizzle)", n, 0, 128 * 16, 128, n * 16, 128, 0, 0, 1, grid=grid)
