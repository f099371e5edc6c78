#!/usr/bin/env python
"""Writes profiles/r02_sass_summary.txt: per-kernel counts of the SASS mnemonics that prove the Blackwell paths
(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = 1-D bulk copy, UTCBAR = tcgen05.commit, HMMA = mma.sync, FFMA, MUFU)
from `cuobjdump -sass` of the in-tree library, plus the `ptxas -v` register / shared-memory / spill lines of a full rebuild.
CPU-only (no GPU needed):  python tests/scripts/sass_summary.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "a3gc_ip_b200", "lib", "liba3gc_b200.so")
CSRC = os.path.join(ROOT, "a3gc_ip_b200", "csrc")
OUT = os.path.join(ROOT, "profiles", "r02_sass_summary.txt")
MNEM = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTCBAR", "SYNCS", "HMMA", "FFMA", "MUFU", "LDS", "STS", "LDG", "STG", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur, total = collections.OrderedDict(), None, collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            total[cur] += 1
            for k in MNEM:
                if op.startswith(k):
                    counts[cur][k] += 1
                    break
    names = demangle(list(counts))
    build = subprocess.run(["make", "-C", CSRC, "-B", "-j", "8"], capture_output=True, text=True)
    ptxas = {}
    cur = None
    for line in (build.stdout + build.stderr).splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = m.group(1)
            ptxas[cur] = []
            continue
        if cur and ("bytes stack frame" in line or "Used " in line):
            ptxas[cur].append(line.strip().replace("ptxas info    : ", ""))
    with open(OUT, "w") as f:
        f.write("# SASS mnemonic counts per kernel of a3gc_ip_b200/lib/liba3gc_b200.so (cuobjdump -sass, sm_100a) and ptxas -v resource lines\n")
        f.write("# UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (1-D TMA), UTCBAR = tcgen05.commit, HMMA = mma.sync\n")
        f.write("# regenerate: python tests/scripts/sass_summary.py\n\n")
        hdr = f"{'kernel':<78} {'instr':>6} " + " ".join(f"{k:>7}" for k in MNEM)
        f.write(hdr + "\n")
        for fn, c in counts.items():
            short = re.sub(r"a3gc::\(anonymous namespace\)::", "", names.get(fn, fn))
            short = re.sub(r"\(.*\)$", "", short)
            f.write(f"{short[:78]:<78} {total[fn]:>6} " + " ".join(f"{c.get(k, 0):>7}" for k in MNEM) + "\n")
        f.write("\n# ptxas -v\n")
        for fn, lines in ptxas.items():
            short = re.sub(r"a3gc::\(anonymous namespace\)::", "", demangle([fn])[fn])
            short = re.sub(r"\(.*\)$", "", short)
            f.write(f"{short}\n")
            for l in lines:
                f.write(f"    {l}\n")
    print("wrote", OUT, "kernels:", len(counts))


if __name__ == "__main__":
    main()
