"""Segmentation fine-tuning — host logic on CPU.

`SegHead` / `FCNHead` / `SegTrainer.seg_step` (hcmoco_b200/segment.py) are built with the fp64 PyTorch statement of the C-ABI
(`kernel_ref.TorchKernels`) in place of `CudaKernels`, so launch order, buffers, strides and accumulate flags are checked exactly
against the oracle (oracle.seg_loss / train_step(seg=...), themselves pinned to the reference's FCNHead by test_oracle_golden.py);
the CUDA kernels are compared on the GPU in tests/test_kernels_gpu.py / test_parity_gpu.py."""
import os
from types import SimpleNamespace

import torch

from engine_check import make_inputs, oracle_state, rel
from kernel_ref import TorchKernels
from oracle import hcmoco_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _head_from_state(K, state, cw):
    from hcmoco_b200.segment import SegHead
    head = SegHead(K, 25, 128, cw)
    head.store.load_state_dict(state)
    return head


def test_seg_head_matches_oracle_and_reference_golden():
    gold = torch.load(os.path.join(GOLD, "seg_head.pt"), weights_only=False)
    K = TorchKernels(dtype=torch.float64)
    cw = gold["class_weights"].double()
    for case in gold["cases"][:3]:
        st_type, tl = case["supervise_type"], case["true_label"]
        sel = torch.nonzero(tl).reshape(-1)
        state = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in case["state"].items()}
        head = _head_from_state(K, state, cw)
        G1, G2 = gold["G1"].double(), gold["G2"].double()
        out2, d1, d2 = head.loss_backward(_nhwc(G1)[sel], _nhwc(G2)[sel], gold["label"][sel], st_type, 10.0)
        # fp64 oracle on the same inputs
        C = {k: v.clone() for k, v in state.items()}
        for k, v in C.items():
            if O.is_param(k):
                v.requires_grad_(True)
        a, b = G1.clone().requires_grad_(True), G2.clone().requires_grad_(True)
        loss, aacc = O.seg_loss(C, a, b, gold["label"], tl, st_type, cw)
        (10.0 * loss).backward()
        assert abs(float(out2[0]) - float(loss)) < 1e-10 and abs(float(out2[1]) - float(aacc)) < 1e-7
        for d, ref in ((d1, a.grad), (d2, b.grad)):
            if ref is None:
                assert d is None
                continue
            assert rel(d, _nhwc(ref)[sel]) < 1e-9
        g = head.store.grads_dict()
        for k in g:
            ref = C[k].grad
            assert rel(g[k], ref) < 1e-8 or float(ref.abs().max()) < 1e-12, k
        sd = head.store.state_dict()
        for k in ("running_mean", "running_var"):
            assert rel(sd["convs.0.norm_name." + k], C["convs.0.norm_name." + k]) < 1e-12
        assert int(sd["convs.0.norm_name.num_batches_tracked"]) == case["nbt"]
        # and against the reference's own numbers (fp32 fixture)
        assert abs(float(out2[0]) - float(case["loss_seg"])) < 2e-5 * float(case["loss_seg"])
        if d1 is not None:
            assert rel(d1, _nhwc(case["d1"])[sel]) < 2e-4
        for k in g:
            assert rel(g[k], case["grads"][k]) < 2e-4 or float(case["grads"][k].abs().max()) < 1e-5, k


def test_fcn_head_module_surface():
    """`build_segmentor(opt)` -> module with the reference's state_dict keys; `classifier(x)` is differentiable."""
    from hcmoco_b200.segment import build_segmentor
    K = TorchKernels(dtype=torch.float64)
    clf = build_segmentor(SimpleNamespace(n_class=25), K)
    lay = O.fcn_layout(25)
    sd = clf.state_dict()
    assert list(sd.keys()) == list(lay.keys())
    assert [tuple(v.shape) for v in sd.values()] == [tuple(s) for s in lay.values()]
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 128, 4, 4, generator=g, dtype=torch.float64, requires_grad=True)
    out = clf(x)
    assert tuple(out.shape) == (2, 25, 16, 16)
    C = {k: v.detach().clone() for k, v in sd.items()}
    C["convs.0.norm_name.num_batches_tracked"] -= 1
    C["convs.0.norm_name.running_mean"].zero_()
    C["convs.0.norm_name.running_var"].fill_(1.0)
    for k, v in C.items():
        if O.is_param(k):
            v.requires_grad_(True)
    x2 = x.detach().clone().requires_grad_(True)
    ref = O.fcn_forward(C, x2)
    assert rel(out, ref) < 1e-10
    w = torch.randn(out.shape, generator=g, dtype=torch.float64)
    (out * w).sum().backward()
    (ref * w).sum().backward()
    assert rel(x.grad, x2.grad) < 1e-9
    for k, p in clf.named_parameters():
        assert rel(p.grad, C[k].grad) < 1e-8 or float(C[k].grad.abs().max()) < 1e-12, k


def test_seg_step_matches_oracle():
    """One fused fine-tuning step (engine forward + head + engine backward + SGD on both stores) against oracle.train_step(seg=...):
    losses, every gradient, parameters and BN buffers after the update — fp64, so the comparison is exact wiring."""
    from hcmoco_b200.api import HCMoCoMem, HCMoCoModel
    from hcmoco_b200.segment import FCNHead, SegTrainer
    cfg = dict(stage=2, width=18, skeleton="mpii", B=3, R=64, K=64, n=500, S=50)
    dt = torch.float64
    K = TorchKernels(dtype=dt)
    layout, P, mom, banks = oracle_state(cfg, dt)
    opt = SimpleNamespace(modal="RGBD2S", arch="HRNet", jigsaw=False, head="linear", pool_method="mean", width=18, linear_feat_map=1,
                          skeleton_meta_name="mpii", in_channel_list=[3, 3], feat_dim=128, mem="bank+jointspri3d", nce_k=cfg["K"],
                          nce_t=0.07, nce_m=0.5, temperature=0.07, pri3d_num_samples_per_image=cfg["S"], modality_missing=1,
                          supervise_type=0, cmc_loss_weights=1, other_loss_weights=1, print_freq=1, n_class=25, cuda_graph=False)
    model = HCMoCoModel(opt, K)
    model.store.load_state_dict(P)
    mem = HCMoCoMem(128, cfg["n"], cfg["K"], 0.07, 0.5, K)
    for i in range(3):
        getattr(mem, "memory_%d" % (i + 1)).copy_(banks[i])
    g = torch.Generator().manual_seed(5)
    cw = (torch.rand(25, generator=g, dtype=dt) + 0.5)
    clf = FCNHead(128, 128, 25, 1, 1, K, cw)
    Cc = {k: v.detach().clone() for k, v in clf.state_dict().items()}
    cmom = O.make_momentum(Cc)
    batch, nce, dense = make_inputs(cfg, 0, dt)
    R = cfg["R"]
    label = torch.randint(0, 25, (cfg["B"], R, R), generator=g)
    label[torch.rand(cfg["B"], R, R, generator=g) < 0.2] = 255
    true_label = torch.tensor([1, 0, 1])
    ref = O.train_step(P, mom, banks, batch, nce, dense, width=18, skeleton="mpii", stage=2, first=True,
                       seg=dict(C=Cc, mom=cmom, label=label, true_label=true_label, supervise_type=0, class_weights=cw))
    tr = SegTrainer(opt)
    tr.injected_dense_idx = dense
    mem.injected_idx = nce
    data = [batch["x"], batch["index"], batch["skeleton"], None, batch["joints_yx"], batch["joints_vis"], batch["use_depth"],
            batch["depth_mask"], None, label, true_label]
    model.attach_memory(mem)
    eng = model.engine_for(cfg["B"], R, mem)
    g_before = {}
    orig_sgd = eng.sgd

    def spy(*a, **k):                       # gradients as they stand when the optimiser runs
        g_before.update(eng.store.grads_dict())
        g_before.update({"clf." + kk: v for kk, v in clf.head.store.grads_dict().items()})
        return orig_sgd(*a, **k)
    eng.sgd = spy
    res = tr.seg_step(model, clf, mem, data, 0.03, 0.9, 1e-4)()
    assert abs(float(res["seg_loss"]) - float(ref["seg_loss"])) < 1e-9
    assert abs(float(res["seg_aacc"]) - float(ref["seg_aacc"])) < 1e-7
    assert abs(float(res["loss"]) - float(ref["loss"])) < 1e-8 * abs(float(ref["loss"]))
    for k, v in ref["grads"].items():
        assert rel(g_before[k], v) < 1e-7 or float(v.abs().max()) < 1e-12, k
    for k, v in ref["seg_grads"].items():
        assert rel(g_before["clf." + k], v) < 1e-7 or float(v.abs().max()) < 1e-12, k
    sd = model.store.state_dict()
    for k in layout:
        if not k.endswith("num_batches_tracked"):
            assert rel(sd[k], P[k]) < 1e-7, k
    sdc = clf.head.store.state_dict()
    for k, v in Cc.items():
        if not k.endswith("num_batches_tracked"):
            assert rel(sdc[k], v) < 1e-7, k
    for m in range(3):
        assert rel(getattr(mem, "memory_%d" % (m + 1)), banks[m]) < 1e-12
