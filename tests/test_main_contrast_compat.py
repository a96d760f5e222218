"""The UNMODIFIED reference launcher (/root/reference/pycontrast/main_contrast.py, `main()` -> `main_worker`, :19-106) executed
over hcmoco_b200/compat: argparse flags as the shipped scripts pass them (scripts/FirstStage/train_ntumpiirgbd2s_hrnet_w18.sh,
scripts/SecondStage/...), build_model / build_mem / ContrastTrainer through the import shims, epoch loop, save, `--resume`,
and a second-stage run started from the first-stage checkpoint with `--pretrain`.

What is NOT the product here (test plumbing, lives in a temp dir): a `datasets/util.py` stub yielding synthetic batch tuples in the
reference's layout (the reference's own data pipeline needs NTU RGB-D / MPII on disk) and a driver that injects the plain-PyTorch
float64 statement of the kernels (tests/kernel_ref.py) for `hcmoco_b200.api._kernels`, records the random draws, and then
`runpy`s main_contrast.py with sys.path = [stub, compat, repo] — i.e. what `PYTHONSAFEPATH=1 PYTHONPATH=compat:repo:.` gives.
The result is compared with the oracle replaying the same batches and draws.

Needs /root/reference (present in the build container only): skipped elsewhere.
"""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

from oracle import hcmoco_oracle as O
from engine_check import rel

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAIN = "/root/reference/pycontrast/main_contrast.py"
pytestmark = [pytest.mark.reference, pytest.mark.skipif(not os.path.exists(MAIN), reason="reference tree not present")]

B, R, N_DATA, NCE_K, S, NB = 2, 64, 300, 64, 50, 2      # NB batches per epoch

STUB_UTIL = '''
import torch
from hcmoco_b200.synthetic import make_batch

class _Dataset:
    def __len__(self):
        return %(n)d

class _Sampler:
    def __init__(self):
        self.epochs = []
    def set_epoch(self, e):
        self.epochs.append(e)

class _Loader(list):
    pass

def build_own_contrast_loader(opt, ngpus_per_node):
    """datasets/util.py:530-597 -> (dataset, loader, sampler); batches in the reference's tuple layout."""
    J = 16 if opt.skeleton_meta_name == "mpii" else 13
    loader = _Loader(make_batch(%(B)d, %(R)d, J, %(n)d, seed=500 + i) for i in range(%(NB)d))
    return _Dataset(), loader, _Sampler()

build_contrast_loader = build_own_contrast_loader
''' % dict(n=N_DATA, B=B, R=R, NB=NB)

DRIVER = '''
import runpy, sys, torch
stub, compat, repo, tests, main, rec_path = sys.argv[1:7]
sys.argv = [main] + sys.argv[7:]
sys.path[:0] = [stub, compat, repo]
sys.path.append(tests)
from kernel_ref import TorchKernels
import hcmoco_b200.api as api
from hcmoco_b200.engine import Engine
api._kernels = lambda kernels=None: kernels if kernels is not None else TorchKernels("cpu", torch.float64)
rec = dict(nce=[], dense=[], init=None)
_build = api.build_model
def build_model(opt, kernels=None):
    m, e = _build(opt, kernels)
    rec["init_built"] = {k: v.clone() for k, v in m.state_dict().items()}
    return m, e
api.build_model = build_model
_draw = api.HCMoCoMem.draw
def draw(self, bsz, y):
    idx = _draw(self, bsz, y)
    rec["nce"].append(idx.clone())
    return idx
api.HCMoCoMem.draw = draw
_dd = Engine.draw_dense
def draw_dense(self, injected=None):
    _dd(self, injected)
    rec["dense"].append(self.dense_idx.clone())
Engine.draw_dense = draw_dense
_train = api.ContrastTrainer.train
def train(self, epoch, loader, model, model_ema, contrast, criterion, optimizer):
    if "start" not in rec:          # state after --pretrain / --resume loading = what the first step starts from
        rec["start"] = {k: v.clone() for k, v in model.state_dict().items()}
        rec["start_banks"] = [getattr(contrast, "memory_%d" % i).clone() for i in (1, 2, 3)]
        rec["start_mom"] = {k: model.store.export(model.store.m, k) for k in model.param_keys}
        rec["first_epoch"] = epoch
    rec.setdefault("lrs", []).append(optimizer.param_groups[0]["lr"])
    out = _train(self, epoch, loader, model, model_ema, contrast, criterion, optimizer)
    rec.setdefault("logs", []).append(out)
    return out
api.ContrastTrainer.train = train
try:
    runpy.run_path(main, run_name="__main__")
finally:
    torch.save(rec, rec_path)
'''


def _run(tmp, tag, extra):
    stub = os.path.join(tmp, "stub")
    os.makedirs(os.path.join(stub, "datasets"), exist_ok=True)
    open(os.path.join(stub, "datasets", "__init__.py"), "w").close()
    with open(os.path.join(stub, "datasets", "util.py"), "w") as f:
        f.write(STUB_UTIL)
    drv = os.path.join(tmp, "driver.py")
    with open(drv, "w") as f:
        f.write(textwrap.dedent(DRIVER))
    rec = os.path.join(tmp, "rec_%s.pt" % tag)
    # the flags of scripts/FirstStage/train_ntumpiirgbd2s_hrnet_w18.sh:15-42 (paths / sizes reduced; no ImageNet weights here)
    args = ("--dataset NTUMPII --data_folder ./data/NTURGBD --train_file_list ./data/x.txt --model_path %s --tb_path %s "
            "--num_workers 0 --learning_rate 0.03 --lr_decay_epochs 40,50,60 --batch_size %d --modal RGBD2S "
            "--in_channel_list 3,3 --nce_k %d --nce_m 0.5 --world-size 1 --rank 0 --multiprocessing-distributed --cosine "
            "--tag t --arch HRNet --width 18 --modality_missing 1 --mpii_root data/mpii/ --pool_method mean --print_freq 1 "
            "--save_freq 1 --pri3d_num_samples_per_image %d " % (os.path.join(tmp, "model"), os.path.join(tmp, "tb"), B, NCE_K, S)
            ).split() + extra
    env = dict(os.environ, OMP_NUM_THREADS="4", RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, drv, stub, os.path.join(REPO, "hcmoco_b200", "compat"), REPO, os.path.join(REPO, "tests"),
                        MAIN, rec] + args, cwd=tmp, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return torch.load(rec, weights_only=False), r.stdout


def _oracle_replay(rec, stage, skeleton, first_global_step):
    """Replay the recorded run with the oracle: same start state, batches, draws and per-epoch learning rates."""
    from hcmoco_b200.synthetic import make_batch
    J = 16 if skeleton == "mpii" else 13
    dt = torch.float64
    P = type(rec["start"])((k, v.clone().to(dt) if v.is_floating_point() else v.clone()) for k, v in rec["start"].items())
    mom = {k: v.clone().to(dt) for k, v in rec["start_mom"].items()}
    banks = [b.clone().to(dt) for b in rec["start_banks"]]
    losses, step = [], 0
    for lr in rec["lrs"]:
        for i in range(NB):
            d = make_batch(B, R, J, N_DATA, seed=500 + i)
            bt = dict(x=d[0].to(dt), index=d[1], skeleton=d[2].to(dt), joints_yx=d[4].to(dt), joints_vis=d[5],
                      use_depth=d[6], depth_mask=d[7].to(dt))
            dense = rec["dense"][step] if stage == 2 else None
            out = O.train_step(P, mom, banks, bt, rec["nce"][step], dense, width=18, skeleton=skeleton, stage=stage, lr=lr,
                               first=(first_global_step + step == 0))
            losses.append(float(out["loss"]))
            step += 1
    return P, banks, losses


def test_unmodified_main_contrast_runs_on_the_engine(tmp_path):
    tmp = str(tmp_path)
    # ---- first stage, one epoch of NB steps, checkpoint
    rec1, out1 = _run(tmp, "s1", "--epochs 1 --method CMCRGBD2S".split())
    assert out1.count("Train: [1]") == NB and "==> Saving..." in out1
    folder = [d for d in os.listdir(os.path.join(tmp, "model")) if d.startswith("CMCRGBD2S_HRNet_RGBD2S")]
    assert len(folder) == 1, folder       # run name derived as options/train_options.py:40-47
    ck_path = os.path.join(tmp, "model", folder[0], "current.pth")
    ck = torch.load(ck_path, map_location="cpu", weights_only=False)
    assert ck["epoch"] == 1 and len(ck["model"]) == 3741 and list(ck["contrast"]) == ["memory_1", "memory_2", "memory_3"]
    assert os.path.exists(os.path.join(tmp, "model", folder[0], "ckpt_epoch_1.pth"))
    P, banks, losses = _oracle_replay(rec1, 1, "mpii", 0)
    assert abs(rec1["logs"][0][0] - sum(losses) / NB) < 1e-8 * abs(losses[0])          # epoch-average loss the trainer returns
    for k, v in ck["model"].items():
        if not k.endswith("num_batches_tracked"):
            assert rel(v, P[k[7:]]) < 1e-7, k
    for i in range(3):
        assert rel(ck["contrast"]["memory_%d" % (i + 1)], banks[i]) < 1e-9
    # ---- --resume: continues at epoch 2 with the saved parameters, momentum and banks
    rec2, out2 = _run(tmp, "s1r", ["--epochs", "2", "--method", "CMCRGBD2S", "--resume", ck_path])
    assert "=> resume successfully" in out2 and rec2["first_epoch"] == 2 and out2.count("Train: [2]") == NB
    for k, v in ck["model"].items():
        if not k.endswith("num_batches_tracked"):
            assert torch.equal(rec2["start"][k[7:]].double(), v.double()), k
    assert float(sum(v.abs().sum() for v in rec2["start_mom"].values())) > 0
    P2, banks2, losses2 = _oracle_replay(rec2, 1, "mpii", NB)
    ck2 = torch.load(ck_path, map_location="cpu", weights_only=False)
    assert ck2["epoch"] == 2
    for k in ("module.encoder1.conv1.weight", "module.encoder2.stage3.1.branches.2.3.conv2.weight", "module.head3.0.bias"):
        assert rel(ck2["model"][k], P2[k[7:]]) < 1e-7, k
    # ---- second stage from the first-stage checkpoint: `--method CMCJointsPri3DRGBD2S` (rejected by the reference's own
    # parser) + `--pretrain` (main_contrast.py:52-67 strips `module.`, keeps unmatched keys, loads the banks)
    rec3, out3 = _run(tmp, "s2", ["--epochs", "1", "--method", "CMCJointsPri3DRGBD2S", "--linear_feat_map", "1", "--pretrain", ck_path])
    assert "Unmatched Keys: encoder1_linear.weight, encoder1_linear.bias, encoder2_linear.weight, encoder2_linear.bias" in out3
    assert torch.equal(rec3["start"]["encoder1.conv1.weight"].double(), ck2["model"]["module.encoder1.conv1.weight"].double())
    assert torch.equal(rec3["start_banks"][2].double(), ck2["contrast"]["memory_3"].double())
    P3, banks3, losses3 = _oracle_replay(rec3, 2, "mpii", 0)
    assert abs(rec3["logs"][0][0] - sum(losses3) / NB) < 1e-8 * abs(losses3[0])
    f3 = [d for d in os.listdir(os.path.join(tmp, "model")) if d.startswith("CMCJointsPri3DRGBD2S")]
    ck3 = torch.load(os.path.join(tmp, "model", f3[0], "current.pth"), map_location="cpu", weights_only=False)
    assert len(ck3["model"]) == 3745
    for k in ("module.encoder1_linear.weight", "module.encoder3.gconv_output.W", "module.encoder2.conv2.weight"):
        assert rel(ck3["model"][k], P3[k[7:]]) < 1e-7, k
