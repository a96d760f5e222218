"""Bring-up check of the fused dense-affinity kernels against the PyTorch statement (tests/kernel_ref.py)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from hcmoco_b200.kernels import CudaKernels  # noqa: E402
from kernel_ref import TorchKernels  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
K, R = CudaKernels(), TorchKernels("cuda")


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30))


for (B, h, S) in ((4, 16, 60), (3, 32, 400), (32, 64, 400), (2, 8, 130)):
    g = torch.Generator().manual_seed(B + S)
    G1 = torch.randn(B, h * h, 128, generator=g).cuda()
    G2 = (0.5 * torch.randn(B, h * h, 128, generator=g)).cuda() + 0.5 * G1
    pix = torch.randint(0, h * h, (B, S), generator=g).cuda()
    kept = (torch.rand(B, generator=g) < 0.7).float().cuda()
    kept[0] = 1
    use_depth = torch.ones(B, dtype=torch.int64).cuda()
    T = 0.07
    work = torch.empty((K.dense_affinity_work_bytes(B, S) + 3) // 4, device="cuda")
    out = {}
    for name, kk in (("cuda", K), ("ref", R)):
        stat, fin = torch.zeros(B, 2, S, 4).cuda(), torch.zeros(8).cuda()
        kk.dense_affinity_fwd(G1, G2, pix, kept, use_depth, B, S, h, 128, 1.0 / T, stat, fin, work)
        out[name] = (stat, fin)
    torch.cuda.synchronize()
    m = kept != 0
    sc, sr = out["cuda"][0][m], out["ref"][0][m]
    print("B=%d h=%d S=%d fwd: lse %.2e Z %.2e wl %.2e hit-mismatch %d/%d  fin %.2e  (loss %s)" % (
        B, h, S, rel(sc[..., 0], sr[..., 0]), rel(sc[..., 1], sr[..., 1]), rel(sc[..., 2], sr[..., 2]),
        int((sc[..., 3] != sr[..., 3]).sum()), sc[..., 3].numel(), rel(out["cuda"][1][:5], out["ref"][1][:5]),
        out["ref"][1][:2].tolist()), flush=True)
    stat, fin = out["ref"]
    d1r, d2r = torch.zeros_like(G1), torch.zeros_like(G2)
    R.dense_affinity_bwd(G1, G2, pix, stat, kept, fin, B, S, h, 128, 1.0 / T, 1.0, 1.0, d1r, d2r, work, 0)
    d1, d2 = torch.zeros_like(G1), torch.zeros_like(G2)
    K.dense_affinity_bwd(G1, G2, pix, stat, kept, fin, B, S, h, 128, 1.0 / T, 1.0, 1.0, d1, d2, work, 0)
    torch.cuda.synchronize()
    print("   bwd: dG1 %.2e dG2 %.2e" % (rel(d1, d1r), rel(d2, d2r)), flush=True)
    if B == 32:
        stat, fin = torch.zeros(B, 2, S, 4).cuda(), torch.zeros(8).cuda()
        d1, d2 = torch.zeros_like(G1), torch.zeros_like(G2)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        for it in range(2):
            ev[0].record()
            for i in range(20):
                K.dense_affinity_fwd(G1, G2, pix, kept, use_depth, B, S, h, 128, 1.0 / T, stat, fin, work)
            ev[1].record()
            for i in range(20):
                K.dense_affinity_bwd(G1, G2, pix, stat, kept, fin, B, S, h, 128, 1.0 / T, 1.0, 1.0, d1, d2, work, 0)
            ev[2].record()
            torch.cuda.synchronize()
        nk = int(kept.sum())
        print("   timing B=32 (kept %d): fwd %.1f us  bwd %.1f us  (algorithmic %.1f MB per pass)" % (
            nk, ev[0].elapsed_time(ev[1]) * 50, ev[1].elapsed_time(ev[2]) * 50, nk * 2 * S * 512 / 1e6))
