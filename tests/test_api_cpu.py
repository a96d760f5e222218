"""The reference-facing surface (hcmoco_b200/api.py) on CPU: the engine's launch programs run with the plain-PyTorch
statement of each kernel (tests/kernel_ref.py, float64), so what is tested here is the host logic of the boundary:
state_dict layout against the reference's own (tests/golden/state_layouts.json), the autograd bridge, the
logits-returning memory module, the trainer's step loop, checkpoints, the optimiser and the launcher flags."""
import argparse
import json
import os

import pytest
import torch
import torch.nn.functional as F

from kernel_ref import TorchKernels
from oracle import hcmoco_oracle as O
from engine_check import make_inputs, oracle_state, rel
from hcmoco_b200 import api
from hcmoco_b200.options import TrainOptions

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CFG = dict(stage=2, width=18, skeleton="coco_reduce", B=3, R=64, K=64, n=300, S=50)


def make_opt(cfg, **kw):
    o = argparse.Namespace(modal="RGBD2S", arch="HRNet", jigsaw=False, head="linear", feat_dim=128, in_channel_list=[3, 3],
                           linear_feat_map=int(cfg["stage"] == 2), width=cfg["width"], pool_method="mean",
                           skeleton_meta_name=cfg["skeleton"], IN_Pretrain=None, depth_Pretrain=None,
                           mem="bank+jointspri3d" if cfg["stage"] == 2 else "bank", nce_k=cfg["K"], nce_t=0.07, nce_m=0.5,
                           temperature=0.07, pri3d_num_samples_per_image=cfg["S"], modality_missing=1, print_freq=1,
                           save_freq=1, learning_rate=0.03, momentum=0.9, weight_decay=1e-4, cosine=True, lr_decay_rate=0.1,
                           lr_decay_epochs=[120, 160], epochs=10, warm=False, resume="", seed=0, cuda_graph=False,
                           local_rank=0, rank=0, world_size=1)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def K64():
    return TorchKernels("cpu", torch.float64)


def test_state_dict_layout_is_the_reference_layout():
    with open(os.path.join(GOLD, "state_layouts.json")) as f:
        lay = json.load(f)
    for name, ref in lay.items():
        w, st, sk = name.split("_", 2)
        cfg = dict(stage=int(st[5:]), width=int(w[1:]), skeleton=sk, K=64, S=50)
        model, ema = api.build_model(make_opt(cfg), kernels=TorchKernels("cpu"))
        assert ema is None
        sd = model.state_dict()
        assert [k for k, _ in ref] == list(sd.keys()), name
        assert [tuple(s) for _, s in ref] == [tuple(v.shape) for v in sd.values()], name
        assert len(list(model.parameters())) == sum(1 for k, _ in ref if not k.endswith(("running_mean", "running_var",
                                                                                        "num_batches_tracked")))
        # encoder1 / encoder2 expose state_dict/load_state_dict (IN_Pretrain path, build_backbone.py:531-560)
        sub = model.encoder1.state_dict()
        model.encoder2.load_state_dict(sub)
        assert torch.equal(model.state_dict()["encoder2.conv1.weight"], sub["conv1.weight"])


def _oracle_losses(out, batch, nce, dense, banks, dtype):
    f1, f2, f3 = torch.chunk(out["f"], 3, 1)
    logits = O.nce_logits(banks, (f1, f2, f3), nce)
    ls, _ = O.nce_losses(logits, batch["use_depth"])
    dl, _ = O.dense_loss(out["linear_merge1"], out["linear_merge2"], batch["depth_mask"], dense, batch["use_depth"])
    jl, _ = O.joint_loss(out["linear_merge1"], out["linear_merge2"], out["feat3"], batch["joints_yx"], batch["joints_vis"],
                         batch["use_depth"])
    sl = O.scl_loss(out["linear_merge1"], out["linear_merge2"], batch["joints_yx"], batch["use_depth"])
    return sum(ls) + sum(dl) + sum(jl) + sl


def test_autograd_bridge_matches_oracle():
    """model(x, s, return_fm=True) -> torch losses on the returned tensors -> loss.backward(): parameter .grad equals
    the oracle's autograd gradient (float64: exact up to rounding)."""
    dt = torch.float64
    layout, P, mom, banks = oracle_state(CFG, dt)
    batch, nce, dense = make_inputs(CFG, 0, dt)
    model, _ = api.build_model(make_opt(CFG), kernels=K64())
    model.store.load_state_dict(P)
    feat1, feat2, feat3, f, aux = model(batch["x"], batch["skeleton"], return_fm=True)
    assert f.shape == (CFG["B"], 384) and aux["linear_merge1"].shape == (CFG["B"], 128, 16, 16)
    assert [tuple(t.shape[1:]) for t in feat1] == [(18, 16, 16), (36, 8, 8), (72, 4, 4), (144, 2, 2)]
    out = dict(f=f, feat3=feat3, linear_merge1=aux["linear_merge1"], linear_merge2=aux["linear_merge2"])
    loss = _oracle_losses(out, batch, nce, dense, banks, dt)
    loss.backward()
    # oracle
    for k, v in P.items():
        if O.is_param(k):
            v.requires_grad_(True)
    P2 = P
    ref_out = O.model_forward(P2, batch["x"], batch["skeleton"], CFG["width"], CFG["skeleton"], CFG["stage"], True)
    ref_loss = _oracle_losses(ref_out, batch, nce, dense, banks, dt)
    ref_loss.backward()
    assert rel(loss, ref_loss) < 1e-10
    sd = dict(model.named_parameters())
    worst = 0.0
    for k, v in P.items():
        if O.is_param(k) and v.grad is not None and float(v.grad.norm()) > 1e-12:
            worst = max(worst, rel(sd[k].grad, v.grad))
    assert worst < 1e-7, worst


def test_memory_module_logits_and_backward():
    dt = torch.float64
    K = K64()
    opt = make_opt(CFG)
    mem = api.build_mem(opt, CFG["n"], kernels=K)
    assert [k for k in mem.state_dict()] == ["memory_1", "memory_2", "memory_3"]
    B, K1 = 3, CFG["K"] + 1
    banks = [getattr(mem, "memory_%d" % i).clone() for i in (1, 2, 3)]
    x = F.normalize(torch.randn(B, 3, 128, dtype=dt), dim=2).reshape(B, 384).requires_grad_(True)
    y = torch.tensor([5, 17, 200])
    idx = torch.randint(0, CFG["n"], (B, K1))
    mem.injected_idx = idx.clone()
    f1, f2, f3 = torch.chunk(x, 3, 1)
    out = mem(f1, f2, f3, y)
    assert len(out) == 7 and out[6].dtype == torch.long and out[0].shape == (B, K1)
    idx[:, 0] = y
    x2 = x.detach().clone().requires_grad_(True)
    ref = O.nce_logits(banks, torch.chunk(x2, 3, 1), idx)
    for a, b in zip(out[:6], ref):
        assert rel(a, b) < 1e-12
    w = torch.randn(6, B, K1, dtype=dt)
    sum((a * w[i]).sum() for i, a in enumerate(out[:6])).backward()
    sum((a * w[i]).sum() for i, a in enumerate(ref)).backward()
    assert rel(x.grad, x2.grad) < 1e-12
    # momentum update happened (rows y), other rows untouched
    for i, bk in enumerate(banks):
        O.bank_update(bk, x2.detach()[:, 128 * i:128 * (i + 1)], y)
        assert rel(getattr(mem, "memory_%d" % (i + 1)), bk) < 1e-12


def test_trainer_loop_checkpoint_and_resume(tmp_path):
    """ContrastTrainer.train over two batches == two oracle steps; save(); a fresh model/memory/optimiser resumes from it."""
    dt = torch.float64
    layout, P, mom, banks = oracle_state(CFG, dt)
    opt = make_opt(CFG, model_folder=str(tmp_path), tb_folder=str(tmp_path))
    K = K64()
    model, _ = api.build_model(opt, kernels=K)
    model.store.load_state_dict(P)
    mem = api.build_mem(opt, CFG["n"], kernels=K)
    for i in range(3):
        getattr(mem, "memory_%d" % (i + 1)).copy_(banks[i])
    trainer = api.build_contrast(opt)
    optimizer = torch.optim.SGD(model.parameters(), lr=0.03, momentum=0.9, weight_decay=1e-4)
    model, _, optimizer = trainer.wrap_up(model, None, optimizer)
    assert isinstance(optimizer, api.FusedSGD)
    refs = []
    for s in range(2):
        batch, nce, dense = make_inputs(CFG, s, dt)
        refs.append(O.train_step(P, mom, banks, batch, nce, dense, width=CFG["width"], skeleton=CFG["skeleton"],
                                 stage=CFG["stage"], first=(s == 0)))
        data = [batch["x"], batch["index"], batch["skeleton"], None, batch["joints_yx"], batch["joints_vis"],
                batch["use_depth"], batch["depth_mask"], None]
        mem.injected_idx = nce.clone()
        trainer.injected_dense_idx = dense
        res = trainer.train_step(model, mem, optimizer, data)()
        assert rel(res["loss"], refs[-1]["loss"]) < 1e-9
    sd = model.state_dict()
    for k in layout:
        if not k.endswith("num_batches_tracked"):
            assert rel(sd[k], P[k]) < 1e-7, k
    for i in range(3):
        assert rel(getattr(mem, "memory_%d" % (i + 1)), banks[i]) < 1e-9
    # checkpoint layout: DDP-style 'module.' keys, bank keys, torch-format optimiser state; transfer_ckpt.py strips
    # 'module.encoder1.' (pycontrast/transfer_ckpt.py:18-23)
    trainer.save(model, None, mem, optimizer, epoch=1)
    ck = torch.load(os.path.join(str(tmp_path), "current.pth"), map_location="cpu", weights_only=False)
    assert set(ck) == {"model", "contrast", "optimizer", "epoch"}
    assert list(ck["model"])[0] == "module.encoder1.conv1.weight" and len(ck["model"]) == len(layout)
    stripped = {k[len("module.encoder1."):]: v for k, v in ck["model"].items() if k.startswith("module.encoder1.")}
    assert len(stripped) == 1830 and "stage4.2.fuse_layers.3.0.2.0.weight" in stripped
    assert list(ck["contrast"]) == ["memory_1", "memory_2", "memory_3"]
    assert "momentum_buffer" in ck["optimizer"]["state"][0]
    opt2 = make_opt(CFG, resume=os.path.join(str(tmp_path), "current.pth"))
    model2, _ = api.build_model(opt2, kernels=K)
    mem2 = api.build_mem(opt2, CFG["n"], kernels=K)
    tr2 = api.ContrastTrainer(opt2)
    _, _, opt2_ = tr2.wrap_up(model2, None, torch.optim.SGD(model2.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4))
    assert tr2.resume_model(model2, None, mem2, opt2_) == 2
    assert rel(model2.store.p, model.store.p) == 0.0 and rel(model2.store.m, model.store.m) == 0.0
    assert opt2_.param_groups[0]["lr"] == optimizer.param_groups[0]["lr"]
    assert torch.equal(mem2.memory_2, mem.memory_2)


def test_lr_schedule_and_flags():
    o = TrainOptions().parse(("--method CMCJointsPri3DRGBD2S --modal RGBD2S --in_channel_list 3,3 --nce_k 16384 --nce_m 0.5 "
                              "--world-size 1 --rank 0 --multiprocessing-distributed --cosine --arch HRNet --width 18 "
                              "--modality_missing 1 --pool_method mean --linear_feat_map 1 --batch_size 224 --epochs 100 "
                              "--learning_rate 0.03 --lr_decay_epochs 40,50,60").split(), make_dirs=False)
    assert o.mem == "bank+jointspri3d" and o.nce_t == 0.07 and o.in_channel_list == [3, 3] and not o.warm
    tr = api.ContrastTrainer(o)
    opt = argparse.Namespace(param_groups=[{"lr": 0.0}])
    for ep in (1, 50, 100):
        tr.adjust_learning_rate(opt, ep)
        assert abs(opt.param_groups[0]["lr"] - O.lr_at_epoch(ep, 100, 0.03)) < 1e-12
    o2 = TrainOptions().parse("--method CMCRGBD2S --batch_size 512 --epochs 100 --cosine".split(), make_dirs=False)
    assert o2.warm and o2.warm_epochs == 5 and o2.mem == "bank"


def test_loss_methods_match_oracle():
    dt = torch.float64
    K = K64()
    B, J, h, S = 3, 13, 16, 40
    g = torch.Generator().manual_seed(0)
    G1, G2 = torch.randn(B, 128, h, h, generator=g, dtype=dt), torch.randn(B, 128, h, h, generator=g, dtype=dt)
    feat3 = torch.randn(B, J, 128, generator=g, dtype=dt)
    batch, _, _ = make_inputs(dict(CFG, B=B), 0, dt)
    dense = torch.randint(0, h * h, (B, S), generator=g)
    tr = api.ContrastTrainer(make_opt(dict(CFG, S=S)))
    l, a = tr._compute_soft_pri3d_loss_accuracy(G1, G2, None, None, use_depth=batch["use_depth"], depth_mask=batch["depth_mask"],
                                                sample_idx=dense, K=K)
    rl, ra = O.dense_loss(G1, G2, batch["depth_mask"], dense, batch["use_depth"])
    assert rel(torch.stack(l), torch.stack(rl)) < 1e-9 and rel(torch.stack(a), torch.stack(ra)) < 1e-9
    l, a = tr._compute_joints_pri3d_loss_accuracy(G1, G2, feat3, None, batch["joints_yx"], batch["joints_vis"],
                                                  use_depth=batch["use_depth"], K=K)
    rl, ra = O.joint_loss(G1, G2, feat3, batch["joints_yx"], batch["joints_vis"], batch["use_depth"])
    assert rel(torch.stack(l), torch.stack(rl)) < 1e-9 and rel(torch.stack(a), torch.stack(ra)) < 1e-9
    l, _ = tr._compute_cross_subject_joints_pri3d_loss(G1, G2, None, None, batch["joints_yx"], batch["joints_vis"],
                                                       use_depth=batch["use_depth"], K=K)
    assert rel(l[0], O.scl_loss(G1, G2, batch["joints_yx"], batch["use_depth"])) < 1e-9


def test_loss_methods_backpropagate_like_the_reference():
    """`sum(losses).backward()` through the three reference-signature loss methods (contrast_trainer.py:642, 744, 830):
    gradients wrt both projection maps and the skeleton features equal the oracle's autograd gradients, also with
    unequal weights on the loss terms."""
    dt = torch.float64
    K = K64()
    B, J, h, S = 3, 13, 16, 40
    g = torch.Generator().manual_seed(1)
    base = [torch.randn(B, 128, h, h, generator=g, dtype=dt), torch.randn(B, 128, h, h, generator=g, dtype=dt),
            torch.randn(B, J, 128, generator=g, dtype=dt)]
    batch, _, _ = make_inputs(dict(CFG, B=B), 0, dt)
    dense = torch.randint(0, h * h, (B, S), generator=g)
    tr = api.ContrastTrainer(make_opt(dict(CFG, S=S)))
    wts = [1.0, 0.25, 2.0, 1.0, 0.5]

    def total(fn_dense, fn_joint, fn_scl, G1, G2, f3):
        dl = fn_dense(G1, G2)
        jl = fn_joint(G1, G2, f3)
        sl = fn_scl(G1, G2)
        return wts[0] * dl[0] + wts[1] * dl[1] + wts[2] * jl[0] + wts[3] * jl[1] + wts[4] * sl

    def run(engine):
        G1, G2, f3 = [t.clone().requires_grad_(True) for t in base]
        if engine:
            t = total(lambda a, b: tr._compute_soft_pri3d_loss_accuracy(a, b, None, None, use_depth=batch["use_depth"],
                                                                        depth_mask=batch["depth_mask"], sample_idx=dense, K=K)[0],
                      lambda a, b, c: tr._compute_joints_pri3d_loss_accuracy(a, b, c, None, batch["joints_yx"], batch["joints_vis"],
                                                                             use_depth=batch["use_depth"], K=K)[0],
                      lambda a, b: tr._compute_cross_subject_joints_pri3d_loss(a, b, None, None, batch["joints_yx"],
                                                                               batch["joints_vis"], use_depth=batch["use_depth"],
                                                                               K=K)[0][0], G1, G2, f3)
        else:
            t = total(lambda a, b: O.dense_loss(a, b, batch["depth_mask"], dense, batch["use_depth"])[0],
                      lambda a, b, c: O.joint_loss(a, b, c, batch["joints_yx"], batch["joints_vis"], batch["use_depth"])[0],
                      lambda a, b: O.scl_loss(a, b, batch["joints_yx"], batch["use_depth"]), G1, G2, f3)
        t.backward()
        return float(t), G1.grad, G2.grad, f3.grad

    e, r = run(True), run(False)
    assert abs(e[0] - r[0]) < 1e-9 * abs(r[0])
    for a, b, name in zip(e[1:], r[1:], ("d/dG1", "d/dG2", "d/dfeat3")):
        assert float(b.abs().sum()) > 0 and rel(a, b) < 1e-8, (name, rel(a, b))


def test_drop_in_loop_zero_grad_backward_step_twice():
    """The documented drop-in loop — optimizer.zero_grad(); loss.backward(); optimizer.step() with the FusedSGD that
    wrap_up returns, and with plain torch.optim.SGD(set_to_none=False) — against the oracle over TWO iterations: a
    gradient buffer aliased between autograd and the engine would double the gradient from the second one on."""
    dt = torch.float64
    cfg = dict(CFG, stage=1)
    for mode in ("fused", "fused_keep", "torch_keep"):
        K = K64()
        layout, P, mom, banks = oracle_state(cfg, dt)
        opt = make_opt(cfg)
        model, _ = api.build_model(opt, kernels=K)
        model.store.load_state_dict(P)
        mem = api.build_mem(opt, cfg["n"], kernels=K)
        for i in range(3):
            getattr(mem, "memory_%d" % (i + 1)).copy_(banks[i])
        tr = api.build_contrast(opt)
        sgd = torch.optim.SGD(model.parameters(), lr=0.03, momentum=0.9, weight_decay=1e-4)
        optimizer = sgd if mode == "torch_keep" else tr.wrap_up(model, None, sgd)[2]
        crit = torch.nn.CrossEntropyLoss()
        for s in range(2):
            batch, nce, dense = make_inputs(cfg, s, dt)
            ref = O.train_step(P, mom, banks, batch, nce, dense, width=cfg["width"], skeleton=cfg["skeleton"], stage=1,
                               first=(s == 0))
            mem.injected_idx = nce.clone()
            f = model(batch["x"], batch["skeleton"])
            f1, f2, f3 = torch.chunk(f, 3, dim=1)
            out = mem(f1, f2, f3, batch["index"])
            sel = batch["use_depth"] == 1
            loss = sum(crit(l[sel] if i <= 3 else l, out[-1][sel] if i <= 3 else out[-1]) for i, l in enumerate(out[:-1]))
            if mode == "fused":
                optimizer.zero_grad()
            else:
                optimizer.zero_grad(set_to_none=False)
            loss.backward()
            g = model.encoder1.conv1.weight.grad
            assert rel(g, ref["grads"]["encoder1.conv1.weight"]) < 1e-7, (mode, s, rel(g, ref["grads"]["encoder1.conv1.weight"]))
            optimizer.step()
            assert rel(loss.detach(), ref["loss"]) < 1e-9, (mode, s)
        sd = model.store.state_dict()
        for k in ("encoder1.conv1.weight", "encoder2.stage4.2.fuse_layers.3.0.2.0.weight", "head3.0.weight"):
            assert rel(sd[k], P[k]) < 1e-8, (mode, k, rel(sd[k], P[k]))


def test_stage1_feat3_gradient_reaches_the_skeleton_encoder():
    """return_fm=True in the first stage: feat3 / avg_feat3 are differentiable outputs (build_backbone.py:296-303); their
    gradient must reach encoder3 (the joint-mean backward accumulates into the seeded slot instead of overwriting it)."""
    dt = torch.float64
    cfg = dict(CFG, stage=1)
    K = K64()
    layout, P, mom, banks = oracle_state(cfg, dt)
    model, _ = api.build_model(make_opt(cfg), kernels=K)
    model.store.load_state_dict(P)
    batch, _, _ = make_inputs(cfg, 0, dt)
    outs = model(batch["x"], batch["skeleton"], return_fm=True)
    feat3, f = outs[2], outs[-1]
    (feat3.square().sum() + f[:, 256:].sum()).backward()
    got = model.encoder3.gconv_output.W.grad.clone()
    for k, v in P.items():
        if O.is_param(k):
            v.requires_grad_(True)
            v.grad = None
    o = O.model_forward(P, batch["x"], batch["skeleton"], cfg["width"], cfg["skeleton"], 1, True)
    (o["feat3"].square().sum() + o["f"][:, 256:].sum()).backward()
    want = P["encoder3.gconv_output.W"].grad
    assert float(want.abs().sum()) > 0 and rel(got, want) < 1e-8, rel(got, want)
