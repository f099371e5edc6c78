#!/usr/bin/env python
"""Stage the UNMODIFIED reference implementation of the hot path under oracle/_ref/ (test infrastructure).

`/root/reference` exists only in the build container; the GPU box sees what travels with the repo snapshot.  The
reference is pure Python (no build system, nothing to compile), so "building" it means copying the few source files
its hot path imports -- byte for byte, never edited -- into the git-ignored directory oracle/_ref/ (ignored by git,
NOT by gpurun, like the built .so files):

    net_aagc.py                      the cells / layers / nets of the path (net_aagc.py:40-695)
    config.py                        `from config import *` at net_aagc.py:3 (joint index sets)
    articulate/**.py                 `import articulate as art` at net_aagc.py:4 (only imported, not called by the nets)
    nira_template_15_norm.pkl        the adjacency prior evaluate_a3gc_tp.py:128-130 loads

`bench.py --impl reference` and bench.py's `cpu_baseline` leg import oracle/_ref/net_aagc.py and time the reference's
own TorchScript cells (jit.ScriptModule, net_aagc.py:68,102,128,177) chained as evaluate_a3gc_tp.py:164-172.
Nothing in a3gc_ip_b200/ may import this directory.

    python oracle/build_ref.py [--check]     (--check: verify an existing copy against /root/reference, no writes)
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = ["net_aagc.py", "config.py", "nira_template_15_norm.pkl",
         "articulate/__init__.py", "articulate/armature.py", "articulate/evaluator.py", "articulate/model.py",
         "articulate/math/__init__.py", "articulate/math/angular.py", "articulate/math/general.py", "articulate/math/spatial.py"]


def available() -> bool:
    return os.path.isfile(os.path.join(DST, "net_aagc.py"))


def stage(check_only: bool = False) -> bool:
    """Copy the files; returns False (and copies nothing) when /root/reference is absent (GPU box)."""
    if not os.path.isdir(REF):
        return False
    ok = True
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        if check_only:
            same = os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)
            ok &= same
            if not same:
                print(f"oracle/_ref/{rel}: differs from {src}")
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
    return ok


def import_reference():
    """Import oracle/_ref/net_aagc.py (the reference, unmodified) as a module; raises if it was never staged."""
    if not available():
        raise RuntimeError("oracle/_ref/ is empty: run `python oracle/build_ref.py` where /root/reference exists")
    import importlib
    import warnings
    if DST not in sys.path:
        sys.path.insert(0, DST)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module("net_aagc")


# ---------------------------------------------------------------------------------------------------------------
# the reference's own three-stage chain (evaluate_a3gc_tp.py:132-145, 164-172) on its own net classes
# ---------------------------------------------------------------------------------------------------------------
REF_NET = {"AAGC": "AAGC_net", "A3GC": "A3GC_net", "AGC": "AGC_net", "GGRU": "G_GRU_net"}
TP_SHAPES = ((12, 3, 256), (15, 3, 64), (15, 9, 128))       # (units_in, units_out, hidden) of stages 1-3


def ref_tp_nets(variant, state_dicts):
    """Three reference nets in eval mode with the given (pose_net.-stripped) state_dicts.  Every parameter is cloned
    first: on CPU the reference aliases all adjacency Parameters of a net to the caller's template
    (`Parameter(adjacency_matrix.t())`, net_aagc.py:56, 88-91, ...), which `.to(cuda)` undoes in its own scripts."""
    import pickle
    import torch
    ref = import_reference()
    with open(os.path.join(DST, "nira_template_15_norm.pkl"), "rb") as f:
        nira = torch.from_numpy(pickle.load(f)).float()
    nets = []
    for (f0, o, h), sd in zip(TP_SHAPES, state_dicts):
        net = getattr(ref, REF_NET[variant])(f0, o, h, nira.clone())
        for p in net.parameters():
            p.data = p.data.clone()
        net.load_state_dict(sd, strict=True)
        nets.append(net.eval())
    return nets


def ref_tp_forward(nets, x):
    """evaluate_a3gc_tp.py:167-171 without the SMPL wrapper: x [B,T,15,12] -> reduced global pose [B,T,15,9]."""
    import torch
    with torch.no_grad():
        y1, _ = nets[0](x, None)
        y2, _ = nets[1](torch.cat((x, y1), dim=-1), None)
        y3, _ = nets[2](torch.cat((x, y2), dim=-1), None)
    return y3


if __name__ == "__main__":
    chk = "--check" in sys.argv
    res = stage(check_only=chk)
    print(("identical" if res else "DIFFERENT / missing") if chk else ("staged -> " + DST if res else "no /root/reference here: nothing staged"))
    sys.exit(0 if res or not chk else 1)
