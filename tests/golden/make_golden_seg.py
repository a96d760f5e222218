"""Golden fixture for the segmentation fine-tuning head, produced by the UNMODIFIED reference classes (build container only):

    python tests/golden/make_golden_seg.py        # writes tests/golden/seg_head.pt

It instantiates `networks.build_linear.build_segmentor` (-> networks/fcn.py FCNHead), loads a seeded state into it, and executes the
lines of `SegTrainer.train_soft_joint_pri3d` that involve it (learning/segment_trainer.py:722-745: boolean selection by true_label,
F.normalize, max over the stacked maps, classifier, criterion_seg = nn.CrossEntropyLoss(ignore_index=255, weight=class_weights) of
main_segmentor.py:76-79, eval_seg_aacc :375-379) with `loss = 10 * loss_seg`, for supervise_type 0, 1, 2 and the no-label branch.
Recorded: inputs, state, loss / aAcc, d(loss)/d(maps), parameter gradients, BN running statistics after the forward."""
import os
import sys
from types import SimpleNamespace

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

import ref_shim  # noqa: E402

WEIGHTS = [1.448, 49.234, 49.483, 48.030, 49.247, 49.492, 48.018, 49.704, 50.052, 49.369, 49.694, 50.090, 49.425, 49.459, 45.846,
           47.156, 45.868, 47.197, 44.167, 42.789, 44.341, 48.632, 48.873, 48.644, 49.004]


def seeded_state(keys_shapes, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, shp in keys_shapes:
        if k.endswith("num_batches_tracked"):
            out[k] = torch.tensor(3, dtype=torch.long)
        elif k.endswith("running_var"):
            out[k] = torch.rand(shp, generator=g) + 0.5
        elif k.endswith("norm_name.weight"):
            out[k] = torch.rand(shp, generator=g) + 0.5
        else:
            out[k] = torch.randn(shp, generator=g) * (0.1 if len(shp) > 1 else 0.3)
    return out


def main():
    assert ref_shim.available()
    ref_shim.install()
    from networks.build_linear import build_segmentor
    from learning.segment_trainer import SegTrainer
    B, h, n_class = 4, 8, 25
    R = 4 * h
    g = torch.Generator().manual_seed(11)
    G1 = torch.randn(B, 128, h, h, generator=g)
    G2 = torch.randn(B, 128, h, h, generator=g)
    label = torch.randint(0, n_class, (B, R, R), generator=g)
    label[torch.rand(B, R, R, generator=g) < 0.15] = 255
    class_weights = torch.from_numpy(np.array(WEIGHTS).astype(np.float32))
    criterion = nn.CrossEntropyLoss(ignore_index=255, weight=class_weights)
    cases = []
    for st_type, true_label in ((0, [1, 0, 1, 1]), (1, [1, 1, 0, 0]), (2, [0, 1, 1, 1]), (0, [0, 0, 0, 0])):
        clf = build_segmentor(SimpleNamespace(n_class=n_class))
        sd0 = seeded_state([(k, tuple(v.shape)) for k, v in clf.state_dict().items()])
        clf.load_state_dict(sd0)
        clf.train()
        tl = torch.tensor(true_label)
        a = G1.clone().requires_grad_(True)
        b = G2.clone().requires_grad_(True)
        aux = {"linear_merge1": a, "linear_merge2": b}
        # ---- learning/segment_trainer.py:722-748, verbatim control flow
        if tl.sum() != 0:
            if st_type == 0:
                lm1 = aux["linear_merge1"][tl.bool()]
                lm2 = aux["linear_merge2"][tl.bool()]
                lm1 = torch.nn.functional.normalize(lm1, dim=1)
                lm2 = torch.nn.functional.normalize(lm2, dim=1)
                mx = torch.max(torch.stack([lm1, lm2]), 0)[0]
                seg_output = clf(mx)
            elif st_type == 1:
                seg_output = clf(torch.nn.functional.normalize(aux["linear_merge1"][tl.bool()], dim=1))
            else:
                seg_output = clf(torch.nn.functional.normalize(aux["linear_merge2"][tl.bool()], dim=1))
            loss_seg = criterion(seg_output, label[tl.bool()])
            aacc = SegTrainer.eval_seg_aacc(None, seg_output, label[tl.bool()])
            loss = loss_seg * 10
        else:
            tmp = clf(aux["linear_merge1"])
            loss_seg = (tmp - tmp).mean()
            aacc = torch.zeros(())
            loss = loss_seg
        loss.backward()
        sd1 = clf.state_dict()
        cases.append(dict(supervise_type=st_type, true_label=tl, state=sd0, loss_seg=loss_seg.detach(), aacc=aacc.detach(),
                          d1=(a.grad if a.grad is not None else torch.zeros_like(a)), d2=(b.grad if b.grad is not None else torch.zeros_like(b)),
                          grads={k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in clf.named_parameters()},
                          running_mean=sd1["convs.0.norm_name.running_mean"].clone(), running_var=sd1["convs.0.norm_name.running_var"].clone(),
                          nbt=int(sd1["convs.0.norm_name.num_batches_tracked"])))
    keys = [(k, list(v.shape)) for k, v in build_segmentor(SimpleNamespace(n_class=n_class)).state_dict().items()]
    torch.save(dict(G1=G1, G2=G2, label=label, class_weights=class_weights, keys=keys, cases=cases), os.path.join(HERE, "seg_head.pt"))
    print("wrote seg_head.pt;", [(c["supervise_type"], float(c["loss_seg"]), float(c["aacc"])) for c in cases])


if __name__ == "__main__":
    main()
