"""bf16-path error of the TP pipelines against the fp32 engine (same weights, same input), with an environment knob
toggled between runs.  Usage: python tests/diag_bf16_error.py "A3GC_TC_RESCALE_BF16=0|A3GC_TC_RESCALE_BF16=1"
Prints rel-L2 / max-abs per stage output for A3GC-TP and AGC-TP on two synthetic inputs."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from a3gc_ip_b200 import synthetic as S


def rel_l2(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main():
    sets = sys.argv[1].split("|") if len(sys.argv) > 1 else [""]
    nira = S.load_nira()
    for variant in ("A3GC", "AGC"):
        for seed in (78, 5):
            x = S.synthetic_input(64, 300, seed=seed).cuda()
            ref, sds = S.build_tp(variant, "cuda", precision="fp32")
            want = [y.float() for y in ref(x)]
            for kv in sets:
                for tok in kv.split():
                    k, v = tok.split("="); os.environ[k] = v
                pipe, _ = S.build_tp(variant, "cuda", precision="bf16", state_dicts=sds)
                got = pipe(x)
                line = " ".join(f"{nm}: {rel_l2(g.float(), w):.3e}/{float((g.float() - w).abs().max()):.3e}"
                                for g, w, nm in zip(got, want, ("y1", "y2", "y3")))
                print(f"{variant} seed {seed} [{kv}] rel-L2/max-abs {line}", flush=True)


if __name__ == "__main__":
    main()
