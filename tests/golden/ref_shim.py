"""Import shim for the UNMODIFIED reference (read-only at /root/reference).

Test infrastructure only: used by tests/golden/make_golden.py (fixture generation, in the
build container) and by the optional `reference-present` tests.  Nothing in hcmoco_b200/
imports this.  The GPU box has no /root/reference; everything here is skipped there.

What it does:
  * stubs `yacs.config.CfgNode`, `pointnet2_cuda`, `tensorboard_logger` (absent deps that the
    hot path never executes);
  * makes `.cuda()` an identity so the reference's unconditional `.cuda()` calls run on CPU;
  * chdir()s to pycontrast/ because official_hrnet.py:485-499 opens its YAML by relative path.
"""
import argparse
import contextlib
import os
import sys
import types

REF_ROOT = os.environ.get("HCMOCO_REFERENCE", "/root/reference")
REF_PY = os.path.join(REF_ROOT, "pycontrast")


def available():
    return os.path.isdir(REF_PY)


class _CfgNode(dict):
    """Just enough of yacs.config.CfgNode for default_config.py + update_from_yaml."""

    def __init__(self, init=None, new_allowed=False, **_):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = _CfgNode(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def defrost(self):
        pass

    def freeze(self):
        pass

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict):
                if k not in self or not isinstance(self[k], dict):
                    self[k] = _CfgNode()
                self[k]._merge(v)
            else:
                self[k] = v

    def merge_from_file(self, path):
        import yaml
        with open(path) as f:
            self._merge(yaml.safe_load(f))

    def merge_from_list(self, lst):
        pass

    def clone(self):
        import copy
        return copy.deepcopy(self)


_installed = False


def install():
    """Idempotently install stubs, chdir and sys.path for the reference."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_PY)
    sys.dont_write_bytecode = True
    yacs = types.ModuleType("yacs")
    yacs_cfg = types.ModuleType("yacs.config")
    yacs_cfg.CfgNode = _CfgNode
    yacs.config = yacs_cfg
    sys.modules.setdefault("yacs", yacs)
    sys.modules.setdefault("yacs.config", yacs_cfg)
    sys.modules.setdefault("pointnet2_cuda", types.ModuleType("pointnet2_cuda"))
    tb = types.ModuleType("tensorboard_logger")

    class Logger:  # noqa: D401
        def __init__(self, logdir=None, flush_secs=2):
            self.values = []

        def log_value(self, name, value, step):
            self.values.append((name, value, step))

    tb.Logger = Logger
    sys.modules.setdefault("tensorboard_logger", tb)

    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self

    os.chdir(REF_PY)
    if REF_PY not in sys.path:
        sys.path.insert(0, REF_PY)
    _installed = True


def make_opt(stage=1, width=18, skeleton="mpii", nce_k=16384, num_samples=400):
    """The opt fields the path reads."""
    return argparse.Namespace(
        jigsaw=False, modal="RGBD2S", arch="HRNet", head="linear", feat_dim=128,
        in_channel_list=[3, 3], linear_feat_map=1 if stage == 2 else 0, width=width,
        pool_method="mean", skeleton_meta_name=skeleton, IN_Pretrain=None, depth_Pretrain=None,
        mem="bank" if stage == 1 else "bank+jointspri3d", nce_k=nce_k, nce_t=0.07, nce_m=0.5,
        temperature=0.07, pri3d_num_samples_per_image=num_samples, modality_missing=1,
        amp=False, gpu=0, rank=0, local_rank=0, world_size=1)


@contextlib.contextmanager
def injected_draws(contrast, nce_idx, dense_idx=None):
    """Replace the two random draws of the path with injected index tensors.

    * `AliasMethod.draw` (memory/alias_multinomial.py:49) -> nce_idx.view(-1)
    * `Tensor.multinomial` (contrast_trainer.py:685) -> dense_idx
    """
    import torch
    old_draw = contrast.multinomial.draw
    old_mn = torch.Tensor.multinomial
    contrast.multinomial.draw = lambda N: nce_idx.reshape(-1).clone()
    if dense_idx is not None:
        torch.Tensor.multinomial = lambda self, num_samples, replacement=False, **k: dense_idx.clone()
    try:
        yield
    finally:
        contrast.multinomial.draw = old_draw
        torch.Tensor.multinomial = old_mn
