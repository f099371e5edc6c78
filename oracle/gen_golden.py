#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (it imports /root/reference, which does not exist on the GPU
box):  ``python oracle/gen_golden.py``.  The fixtures it writes are committed; the tests and the
bench read only those.

What it records
  * data fixtures the path needs at run time: the 15x15 adjacency template
    (nira_template_15_norm.pkl), the channel statistics used by prepare_input
    (data/all_sym_train_stats.pt, data/all_train_stats.pt) and the four shipped graph-net
    checkpoints (trained_models/A3GC, trained_models/G-GRU) re-saved as plain state_dicts;
  * golden outputs of the reference's own classes (net_aagc.py) on seeded inputs:
    single cell steps, full nets with random-init and with trained weights, and the
    three-stage TP chain of evaluate_a3gc_tp.py:164-172 at B=1, T=300 (BASELINE cfg 1);
  * golden output of the reference's own ``prepare_input`` (evaluate_a3gc_tp.py:64-94), executed
    from its source text because the script parses argv at import.

The reference aliases all adjacency Parameters of a net to one buffer on CPU
(``Parameter(adjacency_matrix.t())``, net_aagc.py:56,88-91,...; SURVEY "five things" #3), so every
net is de-aliased (each parameter cloned) before weights are loaded -- what ``.to(cuda)`` does
implicitly in the reference's own scripts.
"""
import ast
import hashlib
import os
import pickle
import sys
import types
import warnings

import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import net_aagc as ref  # noqa: E402  (the reference, unmodified)
from oracle import net_oracle as O  # noqa: E402

NET_CLS = {"AAGC": ref.AAGC_net, "A3GC": ref.A3GC_net, "AGC": ref.AGC_net, "GGRU": ref.G_GRU_net}
CELL_CLS = {"AAGC": ref.AAGC_LSTM_cell, "A3GC": ref.A3GC_LSTM_cell, "AGC": ref.AGC_LSTM_cell, "GGRU": ref.G_GRU_cell}


def load_nira() -> torch.Tensor:
    with open(os.path.join(REF, "nira_template_15_norm.pkl"), "rb") as f:
        return torch.from_numpy(pickle.load(f))           # float64 [15,15]


def dealias(mod: torch.nn.Module) -> torch.nn.Module:
    for p in mod.parameters():
        p.data = p.data.clone()
    return mod


def sd_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().contiguous().numpy().tobytes())
    return h.hexdigest()[:16]


def ref_net(variant, f0, out, hidden, sd):
    net = NET_CLS[variant](f0, out, hidden, load_nira().float())
    dealias(net)
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    # parameter order of the reference == the order our tables claim
    assert [k for k, _ in net.state_dict().items()] == [k for k, _ in O.net_param_shapes(variant, f0, out, hidden)]
    return net.eval()


def flat_states(h):
    out = []
    for s in h:
        out += list(s) if isinstance(s, (tuple, list)) else [s]
    return [t.clone() for t in out]


def run_net_case(name, variant, f0, out, hidden, sd, B, T, xseed, with_h0=False, store_sd=False):
    net = ref_net(variant, f0, out, hidden, sd)
    g = torch.Generator().manual_seed(xseed)
    x = torch.randn(B, T, 15, f0, generator=g)
    x[:, :, [0, 7], :] = 0.0                      # some exactly-zero node rows, as real inputs have
    h0 = None
    if with_h0:
        mk = lambda: 0.5 * torch.randn(B, 15, hidden, generator=g)
        h0 = [mk(), mk()] if variant == "GGRU" else [(mk(), mk()), (mk(), mk())]
    with torch.no_grad():
        h_in = None if h0 is None else [tuple(t.clone() for t in s) if isinstance(s, tuple) else s.clone() for s in h0]
        y, h = net(x, h_in)
    case = dict(name=name, variant=variant, f0=f0, out=out, hidden=hidden, B=B, T=T, x=x, y=y.clone(),
                h_out=flat_states(h), h0=None if h0 is None else flat_states(h0), sd_digest=sd_digest(sd))
    if store_sd:
        case["sd"] = {k: v.clone() for k, v in sd.items()}
    return case


def run_cell_case(variant, f_in, hidden, seed):
    nira = load_nira().float()
    cell = dealias(CELL_CLS[variant](f_in, hidden, nira, activation_fn="tanh"))
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shape in O.cell_param_shapes(variant, f_in, hidden):
        if shape == (15, 15):
            sd[k] = nira.t() + 0.05 * torch.randn(15, 15, generator=g)
        else:
            sd[k] = 0.3 * torch.randn(shape, generator=g)
    cell.load_state_dict(sd, strict=True)
    cell.eval()
    B = 3
    x = torch.randn(B, 15, f_in, generator=g)
    h = 0.5 * torch.randn(B, 15, hidden, generator=g)
    c = 0.5 * torch.randn(B, 15, hidden, generator=g)
    with torch.no_grad():
        if variant == "GGRU":
            o, hn = cell(x, h)
            outs = [o.clone(), hn.clone()]
        else:
            o, (hn, cn) = cell(x, (h, c))
            outs = [o.clone(), hn.clone(), cn.clone()]
    return dict(variant=variant, f_in=f_in, hidden=hidden, sd=sd, x=x, h=h, c=c, outs=outs)


def ref_prepare_input():
    """Execute the reference's own prepare_input source (evaluate_a3gc_tp.py:64-94)."""
    src = open(os.path.join(REF, "evaluate_a3gc_tp.py")).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "prepare_input"][0]
    code = compile(ast.Module(body=[fn], type_ignores=[]), "evaluate_a3gc_tp.py", "exec")
    out = {}
    g = torch.Generator().manual_seed(77)
    oris = [torch.randn(T, 6, 3, 3, generator=g) for T in (7, 4)]
    accs = [3.0 * torch.randn(T, 6, 3, generator=g) for T in (7, 4)]
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        for tag, norm, cda in (("nonorm", False, False), ("norm", True, False), ("norm_cda", True, True)):
            ns = {"torch": torch, "args": types.SimpleNamespace(norm=norm, cda=cda)}
            exec(code, ns)
            res = ns["prepare_input"]([o.view(-1, 54) for o in oris], [a.view(-1, 18) for a in accs], torch.device("cpu"))
            out[tag] = [r.clone() for r in res]
    finally:
        os.chdir(cwd)
    return dict(oris=[o.view(-1, 54) for o in oris], accs=[a.view(-1, 18) for a in accs], outs=out)


def main():
    os.makedirs(OUT, exist_ok=True)
    os.makedirs(os.path.join(OUT, "weights"), exist_ok=True)
    torch.manual_seed(0)
    nira = load_nira()
    torch.save(nira, os.path.join(OUT, "nira_template_15_norm.pt"))
    for nm in ("all_sym_train_stats", "all_train_stats"):
        st = torch.load(os.path.join(REF, "data", nm + ".pt"))
        torch.save({k: {kk: vv.clone() for kk, vv in v.items()} for k, v in st.items() if k in ("acc", "ori")},
                   os.path.join(OUT, nm + ".pt"))

    ckpts = {
        "A3GC_model2": ("trained_models/A3GC/checkpoint_model2_finetuning_9.tar", "A3GC", 15, 3, 64),
        "A3GC_model3": ("trained_models/A3GC/checkpoint_model3_finetuning_8.tar", "A3GC", 15, 9, 128),
        "GGRU_model2": ("trained_models/G-GRU/checkpoint_model2_finetuning_10.tar", "GGRU", 15, 3, 64),
        "GGRU_model3": ("trained_models/G-GRU/checkpoint_model3_finetuning_22.tar", "GGRU", 15, 9, 128),
    }
    trained = {}
    for nm, (path, variant, f0, out, hidden) in ckpts.items():
        ck = torch.load(os.path.join(REF, path), map_location="cpu")
        sd = {k: v.clone().contiguous() for k, v in ck["state_dict"].items()}      # keys keep the 'pose_net.' prefix
        torch.save({"epoch": ck["epoch"], "state_dict": sd}, os.path.join(OUT, "weights", nm + ".pt"))
        trained[nm] = (variant, f0, out, hidden, {k[len("pose_net."):]: v for k, v in sd.items()})

    cases = []
    # (1) every variant, tiny hidden size, everything randomised, sd stored in the fixture
    for i, variant in enumerate(O.VARIANTS):
        sd = O.random_state_dict(variant, 12, 3, 8, nira, seed=100 + i)
        cases.append(run_net_case(f"{variant}_h8_rand", variant, 12, 3, 8, sd, B=3, T=7, xseed=200 + i, store_sd=True))
        sd = O.random_state_dict(variant, 15, 9, 8, nira, seed=110 + i)
        cases.append(run_net_case(f"{variant}_h8_rand_h0", variant, 15, 9, 8, sd, B=2, T=5, xseed=210 + i, with_h0=True, store_sd=True))
    # (2) every variant at the stage shapes the scripts instantiate (H = 64, 128; sd from seed)
    for i, variant in enumerate(O.VARIANTS):
        for (f0, out, hidden) in ((15, 3, 64), (15, 9, 128)):
            sd = O.random_state_dict(variant, f0, out, hidden, nira, seed=300 + i)
            c = run_net_case(f"{variant}_h{hidden}_rand", variant, f0, out, hidden, sd, B=2, T=6, xseed=400 + i)
            c["sd_seed"] = 300 + i
            cases.append(c)
    # stage-1 shape (H=256) for A3GC and G-GRU, short
    for i, variant in enumerate(("A3GC", "GGRU")):
        sd = O.random_state_dict(variant, 12, 3, 256, nira, seed=500 + i)
        c = run_net_case(f"{variant}_h256_rand", variant, 12, 3, 256, sd, B=2, T=4, xseed=600 + i)
        c["sd_seed"] = 500 + i
        cases.append(c)
    # (3) shipped checkpoints
    for nm, (variant, f0, out, hidden, sd) in trained.items():
        c = run_net_case(f"{nm}_trained", variant, f0, out, hidden, sd, B=2, T=12, xseed=700)
        c["weights"] = nm
        cases.append(c)
    torch.save(cases, os.path.join(OUT, "net_cases.pt"))

    cells = [run_cell_case(v, f_in, hidden, 900 + i) for i, v in enumerate(O.VARIANTS) for (f_in, hidden) in ((8, 8), (24, 12))]
    torch.save(cells, os.path.join(OUT, "cell_cases.pt"))

    # (4) BASELINE cfg 1: A3GC-TP and G-GRU-TP chain, B=1, T=300 (stage 1 random seed 0, stages 2-3 trained)
    tp = {}
    for variant, names in (("A3GC", ("A3GC_model2", "A3GC_model3")), ("GGRU", ("GGRU_model2", "GGRU_model3"))):
        sd1 = O.random_state_dict(variant, 12, 3, 256, nira, seed=0)
        nets = [ref_net(variant, 12, 3, 256, sd1), ref_net(*trained[names[0]][:4], trained[names[0]][4]),
                ref_net(*trained[names[1]][:4], trained[names[1]][4])]
        x = O.synthetic_input(1, 300, seed=1234)
        with torch.no_grad():
            y1, _ = nets[0](x)
            y2, _ = nets[1](torch.cat((x, y1.view(1, y1.shape[1], 15, 3)), dim=-1))      # evaluate_a3gc_tp.py:168-169
            y3, _ = nets[2](torch.cat((x, y2.view(1, y2.shape[1], 15, 3)), dim=-1))      # :170-171
        tp[variant] = dict(stage1_seed=0, stage1_digest=sd_digest(sd1), weights=names, x_seed=1234,
                           y1=y1.clone(), y2=y2.clone(), y3=y3.clone())
    torch.save(tp, os.path.join(OUT, "tp_cfg1.pt"))

    torch.save(ref_prepare_input(), os.path.join(OUT, "prepare_input.pt"))

    # pose_loss (net_aagc.py:1077-1087; the class has no super().__init__, call .forward directly)
    g = torch.Generator().manual_seed(5)
    p, t = torch.randn(3, 7, 45, generator=g), torch.randn(3, 7, 45, generator=g)
    pl = ref.pose_loss.__new__(ref.pose_loss)
    pl.loss_weight = None
    torch.save(dict(pred=p, targ=t, loss=pl.forward(p, t)), os.path.join(OUT, "pose_loss.pt"))
    for f in sorted(os.listdir(OUT)):
        pth = os.path.join(OUT, f)
        if os.path.isfile(pth):
            print(f, os.path.getsize(pth))


if __name__ == "__main__":
    main()
