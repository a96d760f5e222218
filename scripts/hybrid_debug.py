"""Debug tool (GPU): run one parity case with selected kernel families replaced by their PyTorch
statement (tests/kernel_ref.py) to locate which CUDA kernels lose accuracy.  Not part of the product."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from kernel_ref import TorchKernels  # noqa: E402
from hcmoco_b200.kernels import CudaKernels  # noqa: E402
from engine_check import run_case  # noqa: E402

GROUPS = {
    "conv_fwd": {"conv2d_fwd", "conv2d_stat_rows"},
    "conv_dgrad": {"conv2d_dgrad"},
    "conv_wgrad": {"conv2d_wgrad"},
    "bn_fwd": {"bn_stats", "bn_finalize", "colstat_rows", "bn_bwd_reduce"},
    "bn_finalize": {"bn_finalize"},
    "bn_apply": {"bn_apply"},
    "bn_bwd": {"bn_bwd_reduce", "bn_bwd_finalize", "bn_bwd_apply", "colstat_rows", "bn_stats"},
    "bn_bwd_apply": {"bn_bwd_apply"},
    "bn_bwd_finalize": {"bn_bwd_finalize"},
    "resample": {"fuse_sum", "upsample_adjoint", "avgpool", "avgpool_bwd", "relu_bwd", "axpy"},
    "gemm": {"gemm"},
}


class Hybrid:
    def __init__(self, torch_names):
        self.c, self.t = CudaKernels(), TorchKernels("cuda")
        self.device, self.dtype, self.launches = "cuda", torch.float32, 0
        self.names = torch_names

    def __getattr__(self, n):
        return getattr(self.t if n in self.names else self.c, n)


cfg = dict(stage=2, width=18, skeleton="coco_reduce", B=3, R=64, K=256, n=1000, S=100)
sets = [[]] + [[g] for g in GROUPS] + [["bn_fwd", "bn_bwd", "bn_apply"], list(GROUPS)]
if len(sys.argv) > 1:
    sets = [s.split("+") if s else [] for s in sys.argv[1:]]
for groups in sets:
    names = set().union(*[GROUPS[g] for g in groups]) if groups else set()
    try:
        rep = run_case(Hybrid(names), cfg, nsteps=2, tol=1.0, gtol=None, gfactor=1e9, resync=True)
        print("torch for %-40s" % "+".join(groups), " | ".join(
            "step%d f %.1e loss %.1e grad %.2e (oracle32 %.2e)" % (s, r["f"], r["loss"], r["grad_global"],
                                                                  r["grad_fp32_oracle_vs_fp64"]) for s, r in rep.items()), flush=True)
    except Exception as ex:  # noqa: BLE001
        print("torch for", groups, "FAILED", repr(ex)[:300], flush=True)
