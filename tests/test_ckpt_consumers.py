"""SURVEY.md section 8(f) rank 2 — the checkpoint consumers, run for real on a checkpoint the ENGINE wrote:

  1. `ContrastTrainer.save` (hcmoco_b200/api.py) writes `current.pth`;
  2. the reference's UNMODIFIED `pycontrast/transfer_ckpt.py` / `transfer_ckpt_depth.py` (:18-23) extract the RGB / depth encoder;
  3. the downstream loaders take the result with their own key / shape rules:
       A2J/hrnet/official_hrnet.py `get_hrnet_w18_backbone(pth)` -> `init_weights` (:456-475)      (depth pose estimation)
       HRNet-Semantic-Segmentation/lib/models/seg_hrnet.py `HighResolutionNet.init_weights` (:456-480)  (human parsing)
     both filter `k in model_dict` and `load_state_dict` the union, so a key or shape mismatch would either raise or silently
     leave the random initialisation in place — the test asserts every backbone tensor arrives bit-exactly.

Needs /root/reference (build container only); the downstream trees are executed in a subprocess under the yacs stub of
tests/golden/ref_shim.py (yacs is not installed here)."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

from kernel_ref import TorchKernels
from hcmoco_b200 import api
from test_api_cpu import make_opt

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = [pytest.mark.reference, pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "A2J")), reason="reference tree not present")]

LOADER = '''
import importlib.util, os, sys, types, torch
import numpy as np
if not hasattr(np, "int"):
    np.int = int                                      # seg_hrnet.py:310 predates numpy 1.24 (version drift, not on the engine's path)
sys.path.insert(0, %(golden)r)
import ref_shim
yacs = types.ModuleType("yacs"); yc = types.ModuleType("yacs.config"); yc.CfgNode = ref_shim._CfgNode; yacs.config = yc
sys.modules["yacs"] = yacs; sys.modules["yacs.config"] = yc
kind, pth, out = sys.argv[1:4]
if kind == "a2j":
    os.chdir(%(ref)r + "/A2J")                       # official_hrnet.py:506 opens ./hrnet/*.yaml relative to the CWD
    sys.path.insert(0, os.getcwd())
    from hrnet.official_hrnet import get_hrnet_w18_backbone
    fresh = get_hrnet_w18_backbone(None)
    model = get_hrnet_w18_backbone(pth)             # -> init_weights(pth): filter `k in model_dict`, load_state_dict
else:
    root = %(ref)r + "/HRNet-Semantic-Segmentation"
    spec = importlib.util.spec_from_file_location("seg_hrnet", root + "/lib/models/seg_hrnet.py")
    seg = importlib.util.module_from_spec(spec); spec.loader.exec_module(seg)
    import yaml
    cfg = ref_shim._CfgNode(yaml.safe_load(open(root + "/experiments/nturgbd_d/config-template.yaml")))
    cfg.MODEL.PRETRAINED = pth
    if "NUM_CLASSES" not in cfg.DATASET:
        cfg.DATASET.NUM_CLASSES = 25
    fresh = seg.HighResolutionNet(cfg)
    model = seg.get_seg_model(cfg)                   # seg_hrnet.py:477-480 -> init_weights(cfg.MODEL.PRETRAINED)
torch.save({"loaded": model.state_dict(), "keys_fresh": list(fresh.state_dict().keys())}, out)
'''


def test_engine_checkpoint_through_transfer_ckpt_into_the_downstream_loaders(tmp_path):
    cfg = dict(stage=2, width=18, skeleton="mpii", B=2, R=64, K=64, n=300, S=50)
    opt = make_opt(cfg, model_folder=str(tmp_path), tb_folder=str(tmp_path))
    K = TorchKernels("cpu", torch.float32)
    model, _ = api.build_model(opt, kernels=K)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():                      # distinct values everywhere (BN statistics included) so that nothing can hide
        model.store.p.copy_(torch.randn(model.store.p.shape, generator=g))
        model.store.bflat.copy_(torch.rand(model.store.bflat.shape, generator=g) + 0.5)
        model.store.nbt.fill_(7)
    mem = api.build_mem(opt, cfg["n"], kernels=K)
    trainer = api.build_contrast(opt)
    _, _, optimizer = trainer.wrap_up(model, None, torch.optim.SGD(model.parameters(), lr=0.03, momentum=0.9, weight_decay=1e-4))
    trainer.save(model, None, mem, optimizer, epoch=1)
    ck = os.path.join(str(tmp_path), "current.pth")
    sd = model.state_dict()
    for script, enc, kind in (("transfer_ckpt.py", "encoder1", "seg"), ("transfer_ckpt_depth.py", "encoder2", "a2j"),
                              ("transfer_ckpt.py", "encoder1", "a2j")):
        out = os.path.join(str(tmp_path), "%s_%s.pth" % (enc, kind))
        r = subprocess.run([sys.executable, os.path.join(REF, "pycontrast", script), "--old_pth", ck, "--new_pth", out],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        bare = torch.load(out, map_location="cpu", weights_only=False)
        want = {k[len(enc) + 1:]: v for k, v in sd.items() if k.startswith(enc + ".")}
        # the whole encoder and — a quirk of the slice test `k[7:15] == 'encoder1'` on a second-stage checkpoint — the 1x1 projection
        # `encoderN_linear.*` as `linear.*`, which both downstream loaders drop (`k in model_dict`); no heads, no other encoder
        assert len(want) == 1830 and set(bare) == set(want) | {"linear.weight", "linear.bias"}
        assert torch.equal(bare["linear.weight"], sd[enc + "_linear.weight"])
        drv = os.path.join(str(tmp_path), "loader_%s.py" % kind)
        with open(drv, "w") as f:
            f.write(textwrap.dedent(LOADER % dict(golden=os.path.join(HERE, "golden"), ref=REF)))
        res = os.path.join(str(tmp_path), "loaded_%s_%s.pt" % (enc, kind))
        r = subprocess.run([sys.executable, drv, kind, out, res], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
        got = torch.load(res, map_location="cpu", weights_only=False)
        loaded = got["loaded"]
        backbone = [k for k in got["keys_fresh"] if not k.startswith("last_layer")]
        assert set(backbone) == set(want), (set(backbone) ^ set(want))     # the downstream backbone has exactly the transferred keys
        for k in backbone:
            assert tuple(loaded[k].shape) == tuple(want[k].shape), k
            assert torch.equal(loaded[k].float(), want[k].float()), (kind, enc, k)
