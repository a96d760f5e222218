"""Generate the golden fixtures by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.pt / *.json

For every case it drives the reference's *own* step loops
(learning/contrast_trainer.py:532 `_train_mem_skeleton3d`, :894 `_train_bank_joints_pri3d_cmc3`)
through `ContrastTrainer.train` with a two-batch in-memory loader, a one-rank gloo group (so
`_global_gather` runs), torch.optim.SGD as in main_contrast.py:78-81, the synthetic state of
synth.py and the synthetic triplets of hcmoco_b200/synthetic.py.  The only interventions are
the ones SURVEY.md Appendix C lists: dependency stubs, `.cuda()` identity, injected random draws,
and the F4 fix (`use_rgb=None` -> all ones) for the SCL loss.

Recorded per step: embeddings, every loss / accuracy, per-parameter gradient L2 norms (a tensor in
named_parameters() order = layout order restricted to parameters), a few full gradients; after
the last step: L2 norm of every state_dict entry (tensor in state_dict order), a few full
tensors and the touched memory-bank rows.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import ref_shim  # noqa: E402
from synth import synthetic_banks, synthetic_state  # noqa: E402
from hcmoco_b200.synthetic import make_batch, make_dense_idx, make_nce_idx  # noqa: E402

CASES = {
    # BASELINE.json configs[0]: first stage, HRNet-w18, bs=2, 224x224
    "c1_stage1_w18_b2_r224": dict(stage=1, width=18, skeleton="mpii", B=2, R=224, K=16384, n=20000, S=400),
    # small second-stage case (all four loss families), MPII skeleton
    "s2_stage2_w18_b4_r128": dict(stage=2, width=18, skeleton="mpii", B=4, R=128, K=1024, n=5000, S=400),
    # COCO-reduced skeleton (J=13), tiny
    "s3_stage2_w18_b3_r64_coco": dict(stage=2, width=18, skeleton="coco_reduce", B=3, R=64, K=256, n=1000, S=100),
    # HRNet-w32 widths, tiny
    "s4_stage2_w32_b2_r64": dict(stage=2, width=32, skeleton="mpii", B=2, R=64, K=256, n=1000, S=100),
}
FULL_KEYS = ["head1.0.weight", "head3.0.bias", "encoder1.conv1.weight", "encoder2.bn1.weight",
             "encoder1.stage4.2.fuse_layers.3.0.2.0.weight", "encoder2.stage3.1.branches.2.3.bn2.bias",
             "encoder3.gconv_layers.2.gconv1.gconv.e", "encoder3.gconv_input.0.gconv.W",
             "encoder1_linear.bias", "encoder2.layer1.0.downsample.0.weight"]
NSTEPS = 2


def run_case(name, cfg):
    from networks.build_backbone import build_model
    from memory.build_memory import build_mem
    from learning.contrast_trainer import ContrastTrainer

    stage, B, R, K, n, S = cfg["stage"], cfg["B"], cfg["R"], cfg["K"], cfg["n"], cfg["S"]
    J = 16 if cfg["skeleton"] == "mpii" else 13
    opt = ref_shim.make_opt(stage, cfg["width"], cfg["skeleton"], K, S)
    opt.warm = False
    opt.print_freq = 1000
    torch.manual_seed(0)
    model, _ = build_model(opt)
    layout = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    state = synthetic_state(layout, seed=0)
    model.load_state_dict(state)
    contrast = build_mem(opt, n)
    banks = synthetic_banks(n, 128, seed=0)
    contrast.memory_1, contrast.memory_2, contrast.memory_3 = [b.clone() for b in banks]

    batches, nce, dense = [], [], []
    h = R // 4
    for s in range(NSTEPS):
        data = make_batch(B, R, J, n, seed=1234 + s)
        batches.append(data)
        nce.append(make_nce_idx(B, K, n, data[1], seed=99 + s))
        dense.append(make_dense_idx(data[7], h, S, seed=7 + s))

    trainer = ContrastTrainer(opt)
    rec = []
    cur = {}

    def wrap(fn_name, key):
        orig = getattr(trainer, fn_name)

        def w(*a, **k):
            if fn_name == "_compute_cross_subject_joints_pri3d_loss" and k.get("use_rgb") is None:
                k["use_rgb"] = torch.ones(B, dtype=torch.long)        # F4
            out = orig(*a, **k)
            cur[key] = out
            return out
        setattr(trainer, fn_name, w)

    wrap("_compute_loss_accuracy", "nce")
    wrap("_compute_soft_pri3d_loss_accuracy", "dense")
    wrap("_compute_joints_pri3d_loss_accuracy", "joint")
    wrap("_compute_cross_subject_joints_pri3d_loss", "scl")

    def fwd_hook(mod, inp, out):
        cur["f"] = (out[3] if isinstance(out, tuple) else out).detach().clone()
        if isinstance(out, tuple):
            cur["feat3"] = out[2].detach().clone()
            cur["lm1"] = out[4]["linear_merge1"].detach().clone()
            cur["lm2"] = out[4]["linear_merge2"].detach().clone()
    model.register_forward_hook(fwd_hook)

    optimizer = torch.optim.SGD(model.parameters(), lr=0.03, momentum=0.9, weight_decay=1e-4)
    orig_step = optimizer.step
    step_no = [0]

    def opt_step(*a, **k):
        s = step_no[0]
        g = {kk: p.grad.detach().clone() for kk, p in model.named_parameters() if p.grad is not None}
        r = dict(f=cur["f"], grad_norm=torch.tensor([float(g[kk].norm()) if kk in g else -1.0
                                                     for kk, _ in model.named_parameters()]),
                 grad_full={kk: g[kk] for kk in FULL_KEYS if kk in g})

        def fl(xs):
            return [float(x) for x in xs]
        r["nce_losses"] = fl(cur["nce"][0])
        r["nce_accs"] = fl(cur["nce"][1])
        if stage == 2:
            r["dense_losses"], r["dense_accs"] = fl(cur["dense"][0]), fl(cur["dense"][1])
            r["joint_losses"], r["joint_accs"] = fl(cur["joint"][0]), fl(cur["joint"][1])
            r["scl_loss"] = float(cur["scl"][0][0])
            r["feat3"] = cur["feat3"]
            r["lm1_slice"] = cur["lm1"][:, ::16, ::5, ::5].clone()
            r["lm2_slice"] = cur["lm2"][:, ::16, ::5, ::5].clone()
            r["lm1_norm"], r["lm2_norm"] = float(cur["lm1"].norm()), float(cur["lm2"].norm())
        rec.append(r)
        step_no[0] += 1
        return orig_step(*a, **k)
    optimizer.step = opt_step

    ce = torch.nn.CrossEntropyLoss()
    criterion = ce if stage == 1 else [ce, [torch.nn.CrossEntropyLoss(), torch.nn.CrossEntropyLoss()]]

    # injected draws: one per step, in call order
    nce_q = [x.reshape(-1).clone() for x in nce]
    contrast.multinomial.draw = lambda N: nce_q.pop(0)
    old_mn = torch.Tensor.multinomial
    dense_q = []
    for s in range(NSTEPS):
        m = batches[s][7][:, ::4, ::4].reshape(B, -1)
        dense_q.append(dense[s][m.sum(-1) > 0].clone())
    torch.Tensor.multinomial = lambda self, num_samples, replacement=False, **k: dense_q.pop(0)
    try:
        trainer.train(1, batches, model, None, contrast, criterion, optimizer)
    finally:
        torch.Tensor.multinomial = old_mn

    sd = model.state_dict()
    final = dict(norm=torch.tensor([float(v.float().norm()) for v in sd.values()]),
                 full={k: sd[k].clone() for k in FULL_KEYS if k in sd})
    for k in ("encoder1.bn1.running_mean", "encoder2.stage4.2.branches.3.3.bn2.running_var",
              "encoder3.gconv_layers.3.gconv2.bn.running_mean", "encoder1.bn1.num_batches_tracked"):
        final["full"][k] = sd[k].clone()
    touched = torch.cat([b[1] for b in batches])
    final["bank_rows"] = [getattr(contrast, "memory_%d" % i)[touched].clone() for i in (1, 2, 3)]
    final["bank_norm"] = [float(getattr(contrast, "memory_%d" % i).norm()) for i in (1, 2, 3)]
    torch.save(dict(cfg=cfg, steps=rec, final=final), os.path.join(HERE, name + ".pt"))
    return layout


def main():
    ref_shim.install()
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29541", rank=0, world_size=1)
    torch.set_num_threads(os.cpu_count())
    layouts = {}
    for name, cfg in CASES.items():
        print("==", name, flush=True)
        layout = run_case(name, cfg)
        layouts["w%d_stage%d_%s" % (cfg["width"], cfg["stage"], cfg["skeleton"])] = \
            [[k, list(s)] for k, s in layout.items()]
    with open(os.path.join(HERE, "state_layouts.json"), "w") as f:
        json.dump(layouts, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
