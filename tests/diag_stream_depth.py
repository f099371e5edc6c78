"""Tuning aid (not a pytest file): L2 -> shared-memory bulk-copy stream rate of a SHALLOW ring (the H = 256 fp32 layer kernel has
three 24 KB slots), i.e. the latency-bound regime: rate = depth * chunk / latency.  python tests/diag_stream_depth.py"""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from a3gc_ip_b200 import _lib

L = _lib.lib()
out = torch.zeros(1, device="cuda")
for span_mb in (8, 64):
    buf = torch.zeros(span_mb << 20, dtype=torch.uint8, device="cuda")
    for grid in (8, 132):
        for chunk in (8192, 16384, 24576):
            for depth in (2, 3, 4, 6, 8):
                for nprod in (1, 2, 3):
                    if depth % nprod:
                        continue
                    for _ in range(2):
                        rc = L.a3gc_tc_stream_bench(buf.data_ptr(), buf.numel(), chunk, depth, 1920, grid, 0, nprod, out.data_ptr(), _lib.stream_ptr(out.device))
                        _lib.check(rc, "a3gc_tc_stream_bench")
                        torch.cuda.synchronize()
                    r = out.item()
                    print(f"span={span_mb:3d}MB grid={grid:3d} chunk={chunk:6d} depth={depth} producers={nprod}: {r:6.1f} B/cycle/CTA  "
                          f"{chunk / r:6.0f} cycles/copy  implied latency {depth * chunk / r:6.0f} cycles", flush=True)
