#!/usr/bin/env python
"""Where a kernel's warp-stall samples fall, from an `ncu --set full --import-source on` report: per window of SASS
instructions (address order follows the source order of the phases) the share of samples, the dominant opcodes and the top
stall reasons, then the single instructions above a threshold.  CPU-only:
    python tests/scripts/ncu_sass_hot.py gpurun_out/<name>.ncu-rep [window=250] [min_share_pct=0.3]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    win = int(sys.argv[2]) if len(sys.argv) > 2 else 250
    thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print(rows[0][1] if len(rows[0]) > 1 else rows[0])
    hdr = rows[1]
    data = [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]
    tot = sum(int(d["# Samples"] or 0) for d in data) or 1
    print(f"{len(data)} SASS instructions, {tot} samples")

    def opcode(src):
        parts = src.split()
        if not parts:
            return ""
        return (parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]).split(".")[0]

    for i in range(0, len(data), win):
        seg = data[i:i + win]
        s = sum(int(d["# Samples"] or 0) for d in seg)
        if s < tot * 0.002:
            continue
        ops, st = {}, {}
        for d in seg:
            ops[opcode(d["Source"])] = ops.get(opcode(d["Source"]), 0) + 1
            for k, v in d.items():
                if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0"):
                    st[k] = st.get(k, 0) + int(v)
        top = ", ".join(f"{k} {v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:5])
        ts = ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{i:6d}-{i + win:6d}: {100 * s / tot:5.1f}%  [{top}]  [{ts}]")
    print(f"--- instructions with >= {thr} % of the samples")
    for i, d in enumerate(data):
        s = int(d["# Samples"] or 0)
        if s >= tot * thr / 100:
            print(f"{i:6d} {100 * s / tot:5.2f}%  {d['Source'][:100]}")


if __name__ == "__main__":
    main()
