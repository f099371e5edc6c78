"""Tuning aid (not a pytest file): L2 -> shared-memory bulk-copy stream rate vs number of active SMs, copy size, ring depth,
number of producer threads and 2-CTA multicast.  python tests/diag_stream_rate.py"""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from a3gc_ip_b200 import _lib

L = _lib.lib()
out = torch.zeros(1, device="cuda")
buf = torch.zeros(16 << 20, dtype=torch.uint8, device="cuda")        # 16 MB: L2 resident


def run(chunk, depth, grid, mcast=0, nprod=1, iters=1920):
    for _ in range(2):
        rc = L.a3gc_tc_stream_bench(buf.data_ptr(), buf.numel(), chunk, depth, iters, grid, mcast, nprod, out.data_ptr(), _lib.stream_ptr(out.device))
        _lib.check(rc, "a3gc_tc_stream_bench")
        torch.cuda.synchronize()
    r = out.item()
    print(f"chunk={chunk:6d} depth={depth} grid={grid:3d} mcast={mcast} producers={nprod}: {r:6.1f} B/cycle/CTA  {chunk / r:7.0f} cycles/copy  ({r * grid:8.0f} B/cycle chip-wide)")


for grid in (8, 148):
    for chunk, depth in ((4096, 8), (16384, 8), (24576, 8), (49152, 4), (65536, 3)):
        run(chunk, depth, grid)
for grid in (8, 132, 148):
    for nprod in (1, 2, 4):
        run(16384, 8, grid, nprod=nprod)
        run(24576, 8, grid, nprod=nprod)
run(16384, 4, 148, mcast=1)
