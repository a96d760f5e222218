"""Pin the CPU oracle (oracle/hcmoco_oracle.py) against fixtures produced by the reference itself
(tests/golden/make_golden.py).  Tolerances: the oracle and the reference run the same ATen fp32
kernels in a different op order, so agreement is ~1e-5; the bar written here is 2e-4 relative on
losses / embeddings and 2e-3 on per-parameter gradient norms (tiny-norm gradients through 300
train-mode BNs amplify rounding)."""
import json
import os

import pytest
import torch

from oracle import hcmoco_oracle as O
from synth import synthetic_banks, synthetic_state
from hcmoco_b200.synthetic import make_batch, make_dense_idx, make_nce_idx

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["c1_stage1_w18_b2_r224", "s2_stage2_w18_b4_r128", "s3_stage2_w18_b3_r64_coco", "s4_stage2_w32_b2_r64"]


def _rel(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def test_layouts_match_reference():
    with open(os.path.join(GOLD, "state_layouts.json")) as f:
        lay = json.load(f)
    for name, ref in lay.items():
        w, st, sk = name.split("_", 2)
        mine = O.model_layout(int(w[1:]), int(st[5:]), sk)
        assert [k for k, _ in ref] == list(mine.keys()), name
        assert [tuple(s) for _, s in ref] == [tuple(s) for s in mine.values()], name


def run_oracle_case(name):
    gold = torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)
    cfg = gold["cfg"]
    J = 16 if cfg["skeleton"] == "mpii" else 13
    layout = O.model_layout(cfg["width"], cfg["stage"], cfg["skeleton"])
    P = synthetic_state(layout, 0)
    mom = O.make_momentum(P)
    banks = synthetic_banks(cfg["n"], 128, 0)
    outs = []
    batches = []
    for s in range(len(gold["steps"])):
        d = make_batch(cfg["B"], cfg["R"], J, cfg["n"], seed=1234 + s)
        batches.append(d)
        batch = dict(x=d[0], index=d[1], skeleton=d[2], joints_yx=d[4], joints_vis=d[5], use_depth=d[6],
                     depth_mask=d[7])
        nce = make_nce_idx(cfg["B"], cfg["K"], cfg["n"], d[1], seed=99 + s)
        dense = make_dense_idx(d[7], cfg["R"] // 4, cfg["S"], seed=7 + s)
        outs.append(O.train_step(P, mom, banks, batch, nce, dense, width=cfg["width"], skeleton=cfg["skeleton"],
                                 stage=cfg["stage"], first=(s == 0)))
    return gold, cfg, layout, P, banks, outs, batches


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    gold, cfg, layout, P, banks, outs, batches = run_oracle_case(name)
    pkeys = [k for k in layout if O.is_param(k)]
    for s, (g, o) in enumerate(zip(gold["steps"], outs)):
        assert _rel(o["f"], g["f"]) < 2e-4
        assert _rel(o["nce_losses"], g["nce_losses"]) < 2e-4
        assert _rel(o["nce_accs"], g["nce_accs"]) < 1e-6 or g["nce_accs"] == [float(a) for a in o["nce_accs"]]
        if cfg["stage"] == 2:
            assert _rel(o["dense_losses"], g["dense_losses"]) < 2e-4
            assert _rel(o["joint_losses"], g["joint_losses"]) < 2e-4
            assert _rel(o["scl_loss"], g["scl_loss"]) < 2e-4
            assert _rel(o["dense_accs"], g["dense_accs"]) < 1e-5
            assert _rel(o["joint_accs"], g["joint_accs"]) < 1e-5
            assert _rel(o["feat3"], g["feat3"]) < 2e-4
            assert _rel(o["linear_merge1"][:, ::16, ::5, ::5], g["lm1_slice"]) < 2e-4
            assert _rel(o["linear_merge2"][:, ::16, ::5, ::5], g["lm2_slice"]) < 2e-4
        gn = torch.tensor([float(o["grads"][k].norm()) if k in o["grads"] else -1.0 for k in pkeys])
        assert gn.shape == g["grad_norm"].shape
        assert _rel(gn, g["grad_norm"]) < 2e-3
        for k, v in g["grad_full"].items():
            assert _rel(o["grads"][k], v) < 5e-3, (s, k)
    fin = gold["final"]
    norms = torch.tensor([float(P[k].float().norm()) for k in layout])
    assert _rel(norms, fin["norm"]) < 1e-4
    for k, v in fin["full"].items():
        assert _rel(P[k], v) < 1e-3, k
    touched = torch.cat([b[1] for b in batches])
    for bk, rows, nrm in zip(banks, fin["bank_rows"], fin["bank_norm"]):
        assert _rel(bk[touched], rows) < 2e-4
        assert abs(float(bk.norm()) - nrm) / nrm < 1e-5


def test_oracle_seg_head_matches_reference():
    """oracle.seg_loss / fcn_forward (restating networks/fcn.py + learning/segment_trainer.py:722-748) against the fixture written by
    tests/golden/make_golden_seg.py from the reference's own FCNHead, nn.CrossEntropyLoss and eval_seg_aacc."""
    gold = torch.load(os.path.join(GOLD, "seg_head.pt"), weights_only=False)
    lay = O.fcn_layout(25)
    assert [k for k, _ in gold["keys"]] == list(lay.keys())
    assert [tuple(s) for _, s in gold["keys"]] == [tuple(s) for s in lay.values()]
    for case in gold["cases"]:
        C = {k: v.clone() for k, v in case["state"].items()}
        for k, v in C.items():
            if O.is_param(k):
                v.requires_grad_(True)
        G1, G2 = gold["G1"].clone().requires_grad_(True), gold["G2"].clone().requires_grad_(True)
        loss, aacc = O.seg_loss(C, G1, G2, gold["label"], case["true_label"], case["supervise_type"], gold["class_weights"])
        (10.0 * loss).backward()
        assert abs(float(loss) - float(case["loss_seg"])) <= 2e-5 * max(1.0, abs(float(case["loss_seg"])))
        assert abs(float(aacc) - float(case["aacc"])) < 1e-6
        for g, ref in ((G1.grad, case["d1"]), (G2.grad, case["d2"])):
            g = torch.zeros_like(ref) if g is None else g
            assert _rel(g, ref) < 2e-4 or float(ref.abs().max()) == 0.0
        for k, ref in case["grads"].items():
            g = C[k].grad if C[k].grad is not None else torch.zeros_like(C[k])
            # (the conv bias in front of a train-mode BatchNorm has an exactly-zero gradient: only rounding noise on both sides)
            assert _rel(g, ref) < 2e-4 or float(ref.abs().max()) < 1e-5, k
        assert _rel(C["convs.0.norm_name.running_mean"], case["running_mean"]) < 1e-5
        assert _rel(C["convs.0.norm_name.running_var"], case["running_var"]) < 1e-5
        assert int(C["convs.0.norm_name.num_batches_tracked"]) == case["nbt"]
