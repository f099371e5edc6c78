"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden/, made by
oracle/gen_golden.py).  fp32 oracle vs fp32 reference: tolerance 2e-6 rel-L2 (same op order, the
only differences are ATen blocking choices); the fp64 oracle is within 2e-6 as well, which bounds
the reference's own fp32 noise floor (2.8e-7)."""
import os

import pytest
import torch

from conftest import load_golden, rel_l2, max_rel, GOLDEN
from oracle import net_oracle as O

NET_CASES = load_golden("net_cases.pt")
CELL_CASES = load_golden("cell_cases.pt")
TOL = 2e-6


def case_sd(case, nira):
    if "sd" in case:
        return case["sd"]
    if "weights" in case:
        ck = load_golden(os.path.join("weights", case["weights"] + ".pt"))
        return {k[len("pose_net."):]: v for k, v in ck["state_dict"].items()}
    return O.random_state_dict(case["variant"], case["f0"], case["out"], case["hidden"], nira, seed=case["sd_seed"])


def unflatten_h(variant, flat):
    if flat is None:
        return None
    if variant == "GGRU":
        return [flat[0], flat[1]]
    return [(flat[0], flat[1]), (flat[2], flat[3])]


@pytest.mark.parametrize("case", NET_CASES, ids=[c["name"] for c in NET_CASES])
def test_net_matches_reference(case, nira):
    sd = case_sd(case, nira)
    with torch.no_grad():
        y, h = O.net_forward(case["variant"], case["x"], sd, unflatten_h(case["variant"], case["h0"]))
    assert y.shape == case["y"].shape
    assert rel_l2(y, case["y"]) < TOL and max_rel(y, case["y"]) < 5 * TOL
    flat = []
    for s in h:
        flat += list(s) if isinstance(s, tuple) else [s]
    for a, b in zip(flat, case["h_out"]):
        assert rel_l2(a, b) < TOL


@pytest.mark.parametrize("case", [c for c in NET_CASES if c["hidden"] <= 64], ids=lambda c: c["name"])
def test_net_fp64_truth(case, nira):
    sd = O.cast_sd(case_sd(case, nira), torch.float64)
    h0 = unflatten_h(case["variant"], None if case["h0"] is None else [t.double() for t in case["h0"]])
    with torch.no_grad():
        y, _ = O.net_forward(case["variant"], case["x"].double(), sd, h0)
    assert rel_l2(y, case["y"]) < TOL


@pytest.mark.parametrize("case", CELL_CASES, ids=[f'{c["variant"]}_{c["f_in"]}_{c["hidden"]}' for c in CELL_CASES])
def test_cell_matches_reference(case):
    sd = {"cell." + k: v for k, v in case["sd"].items()}
    with torch.no_grad():
        if case["variant"] == "GGRU":
            outs = list(O.cell_ggru(case["x"], case["h"], sd, "cell."))
        else:
            o, (h, c) = O.cell_lstm(case["variant"], case["x"], (case["h"], case["c"]), sd, "cell.")
            outs = [o, h, c]
    for a, b in zip(outs, case["outs"]):
        assert rel_l2(a, b) < TOL


def test_param_tables_match_checkpoints():
    for nm, variant, f0, out, hidden, total in (("A3GC_model2", "A3GC", 15, 3, 64, 220049), ("A3GC_model3", "A3GC", 15, 9, 128, 863511),
                                                ("GGRU_model2", "GGRU", 15, 3, 64, 143693), ("GGRU_model3", "GGRU", 15, 9, 128, 565203)):
        ck = load_golden(os.path.join("weights", nm + ".pt"))["state_dict"]
        tbl = O.net_param_shapes(variant, f0, out, hidden)
        assert [k for k in ck] == ["pose_net." + k for k, _ in tbl]
        assert all(tuple(ck["pose_net." + k].shape) == s for k, s in tbl)
        assert sum(v.numel() for v in ck.values()) == total           # SURVEY.md section 8b net totals


def test_tp_chain_cfg1(nira):
    """BASELINE cfg 1: A3GC-TP / G-GRU-TP, B=1, T=300, stage 1 random (seed 0), stages 2-3 trained."""
    tp = load_golden("tp_cfg1.pt")
    for variant, g in tp.items():
        sds = [O.random_state_dict(variant, 12, 3, 256, nira, seed=g["stage1_seed"])]
        for nm in g["weights"]:
            ck = load_golden(os.path.join("weights", nm + ".pt"))
            sds.append({k[len("pose_net."):]: v for k, v in ck["state_dict"].items()})
        x = O.synthetic_input(1, 300, seed=g["x_seed"])
        with torch.no_grad():
            y1, y2, y3 = O.tp_forward(variant, x, sds)
        for a, b in ((y1, g["y1"]), (y2, g["y2"]), (y3, g["y3"])):
            assert rel_l2(a, b) < 5e-6, variant


def test_prepare_input_bit_exact():
    """Index handling (IMU drop, node scatter [3,4,13,14,10]) and normalisation: bit-exact."""
    g = load_golden("prepare_input.pt")
    for tag, stats_name in (("nonorm", None), ("norm", "all_train_stats.pt"), ("norm_cda", "all_sym_train_stats.pt")):
        stats = None if stats_name is None else load_golden(stats_name)
        for ori, acc, want in zip(g["oris"], g["accs"], g["outs"][tag]):
            got = O.prepare_input(ori, acc, stats).unsqueeze(0)
            assert got.shape == want.shape and torch.equal(got, want), tag
            zero_nodes = [n for n in range(15) if n not in O.INPUT_JOINTS]
            assert torch.count_nonzero(got[:, :, zero_nodes]) == 0


def test_pose_loss():
    g = load_golden("pose_loss.pt")
    assert torch.equal(O.pose_loss(g["pred"], g["targ"]), g["loss"])


def test_reverse_direction_final_state_is_t0(nira):
    """Property (SURVEY 'five things' #4): the reverse layer's final state is the state after t=0."""
    sd = O.random_state_dict("A3GC", 12, 3, 8, nira, seed=1)
    x = torch.randn(5, 2, 15, 8, generator=torch.Generator().manual_seed(2))
    z = torch.zeros(2, 15, 8)
    with torch.no_grad():
        out, (h, c) = O.layer_forward("A3GC", x, (z, z), sd, "rnn1.directions.1.", reverse=True)
        # re-run last processed step (t=0) by hand from the state after t=1
        _, st = O.layer_forward("A3GC", x[1:], (z, z), sd, "rnn1.directions.1.", reverse=True)
        o0, (h0, c0) = O.cell_lstm("A3GC", x[0], st, sd, "rnn1.directions.1.cell.")
    assert torch.allclose(h, h0) and torch.allclose(c, c0) and torch.allclose(out[0], o0)


def test_oracle_ik_post_step_matches_reference_golden():
    """Reduced-global -> full-local pose (net_aagc.py:788-800) against the reference's own functions (oracle/gen_golden_ik.py)."""
    g = load_golden("ik_cases.pt")
    assert g["parent"] == O.SMPL_PARENT and g["reduced"] == O.JOINT_REDUCED and g["ignored"] == O.JOINT_IGNORED
    assert sorted(g["reduced"] + g["ignored"]) == list(range(24))
    y9 = O.reduced_global_to_full_local(g["x9"], 9)
    y6 = O.reduced_global_to_full_local(g["x6"], 6)
    assert torch.equal(y9, g["y9"])
    assert (y6 - g["y6"]).abs().max() <= 1e-6
