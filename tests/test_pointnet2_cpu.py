"""PointNet++ primitives — CPU side.

  * the numpy restatement (oracle/pn2_oracle.py) against brute-force definitions written here in plain Python loops (small clouds);
    its bit-level pin against the reference's CUDA kernels happens on the GPU box (tests/test_pointnet2_gpu.py);
  * `hcmoco_b200.pointnet2_utils` (the reference's Python surface over the hcm_pn2_* entry points) with the reference executor:
    shapes, dtypes and gradients of the six functions, and — when /root/reference is present — the reference's UNMODIFIED
    Pointnet2MSG (networks/pointnet2_msg.py + pointnet2_modules.py) running forward and backward on these functions."""
import os
import sys

import numpy as np
import pytest
import torch

from kernel_ref import TorchKernels
from oracle import pn2_oracle as PO


def _cloud(B, N, seed):
    rng = np.random.default_rng(seed)
    x = rng.integers(-16, 17, (B, N, 3)).astype(np.float32) / 16.0
    x[:, N // 2:] = x[:, :N - N // 2]                     # exact duplicates
    return x


def test_oracle_against_brute_force_definitions():
    B, N, M = 2, 48, 12
    xyz = _cloud(B, N, 0)
    # farthest point sampling: greedy max-min; among equal distances the reference's layout rule (T = 32 threads here)
    idx = PO.furthest_point_sampling(xyz, M)
    T = 32

    def brev(v, bits=5):
        return int(format(v, "0%db" % bits)[::-1], 2)
    for b in range(B):
        md = [1e10] * N
        old = 0
        assert idx[b, 0] == 0
        for j in range(1, M):
            for k in range(N):
                d = float(PO.sqdist(xyz[b, k][None], xyz[b, old][None])[0])
                md[k] = min(md[k], d)
            best = max(md)
            cands = [k for k in range(N) if md[k] == best]
            old = min(cands, key=lambda k: (brev(k % T), k))
            assert idx[b, j] == old
    # ball query / three_nn / gathers
    new_xyz = xyz[:, :5].copy()
    bq = PO.ball_query(0.5, 6, xyz, new_xyz)
    d2, nn = PO.three_nn(new_xyz, xyz)
    for b in range(B):
        for q in range(5):
            hits = [k for k in range(N) if float(PO.sqdist(new_xyz[b, q][None], xyz[b, k][None])[0]) < np.float32(0.5) * np.float32(0.5)]
            exp = (hits[:6] + [hits[0]] * 6)[:6] if hits else [0] * 6
            assert list(bq[b, q]) == exp
            ds = sorted((float(PO.sqdist(new_xyz[b, q][None], xyz[b, k][None])[0]), k) for k in range(N))[:3]
            assert [k for _, k in ds] == list(nn[b, q]) and [d for d, _ in ds] == [float(v) for v in d2[b, q]]
    f = np.random.default_rng(1).standard_normal((B, 3, N)).astype(np.float32)
    g = PO.group_points(f, bq)
    assert g.shape == (B, 3, 5, 6) and g[1, 2, 3, 4] == f[1, 2, bq[1, 3, 4]]
    assert PO.gather_points(f, idx)[0, 1, 7] == f[0, 1, idx[0, 7]]


def test_python_surface_shapes_and_gradients():
    import hcmoco_b200.pointnet2_utils as U
    U.set_kernels(TorchKernels())
    try:
        B, N, M = 2, 40, 8
        xyz = torch.from_numpy(_cloud(B, N, 3))
        idx = U.furthest_point_sample(xyz, M)
        assert idx.dtype == torch.int32 and tuple(idx.shape) == (B, M)
        feats = torch.randn(B, 5, N, requires_grad=True)
        got = U.gather_operation(feats, idx)
        ref = torch.gather(feats, 2, idx.long()[:, None, :].expand(B, 5, M))
        assert torch.equal(got, ref)
        w = torch.randn_like(got)
        (g1,) = torch.autograd.grad((got * w).sum(), feats)
        (g2,) = torch.autograd.grad((ref * w).sum(), feats)
        assert torch.allclose(g1, g2, atol=1e-6)
        new_xyz = torch.gather(xyz, 1, idx.long()[..., None].expand(B, M, 3)).contiguous()
        bq = U.ball_query(0.6, 4, xyz, new_xyz)
        grp = U.grouping_operation(feats, bq)
        refg = torch.gather(feats, 2, bq.long().reshape(B, 1, -1).expand(B, 5, M * 4)).reshape(B, 5, M, 4)
        assert torch.equal(grp, refg)
        (g1,) = torch.autograd.grad(grp.square().sum(), feats)
        (g2,) = torch.autograd.grad(refg.square().sum(), feats)
        assert torch.allclose(g1, g2, atol=1e-5)
        dist, nn = U.three_nn(xyz, new_xyz)
        assert tuple(dist.shape) == (B, N, 3) and nn.dtype == torch.int32
        known_f = torch.randn(B, 5, M, requires_grad=True)
        wgt = torch.rand(B, N, 3)
        out = U.three_interpolate(known_f, nn, wgt)
        refo = sum(torch.gather(known_f, 2, nn[:, :, j].long()[:, None, :].expand(B, 5, N)) * wgt[:, None, :, j] for j in range(3))
        assert torch.allclose(out, refo, atol=1e-6)
        (g1,) = torch.autograd.grad(out.sum(), known_f)
        (g2,) = torch.autograd.grad(refo.sum(), known_f)
        assert torch.allclose(g1, g2, atol=1e-5)
        qg = U.QueryAndGroup(0.6, 4)(xyz, new_xyz, feats)
        assert tuple(qg.shape) == (B, 3 + 5, M, 4)
    finally:
        U.set_kernels(None)


def test_reference_pointnet2_msg_runs_on_the_replacement_functions():
    """The reference's own Pointnet2MSG (unmodified modules from /root/reference) with its `pointnet2_utils` functions swapped
    for ours: the FFI consumer is a drop-in.  Skipped where the reference tree is absent (the GPU box)."""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    ref_shim.install()
    import hcmoco_b200.pointnet2_utils as U
    from networks.pointnet2 import pointnet2_utils as RU
    from networks.pointnet2 import pointnet2_modules as RM
    from networks import pointnet2_msg as MSG
    U.set_kernels(TorchKernels())
    saved = {n: getattr(RU, n) for n in ("furthest_point_sample", "gather_operation", "three_nn", "three_interpolate",
                                         "grouping_operation", "ball_query")}
    saved_cfg = (MSG.NPOINTS, MSG.RADIUS)
    try:
        for n in saved:
            setattr(RU, n, getattr(U, n))
        MSG.NPOINTS, MSG.RADIUS = [64, 32, 16, 8], [[0.2, 0.4], [0.4, 0.6], [0.6, 0.9], [0.9, 1.5]]
        torch.manual_seed(0)
        net = MSG.Pointnet2MSG(input_channels=0)
        net.train()
        pts = torch.from_numpy(_cloud(2, 64, 5)).requires_grad_(False)
        out = net(pts)
        assert tuple(out.shape) == (2, 128, 64) and torch.isfinite(out).all()
        out.square().mean().backward()
        gsum = sum(float(p.grad.abs().sum()) for p in net.parameters() if p.grad is not None)
        assert gsum > 0 and np.isfinite(gsum)
    finally:
        for n, f in saved.items():
            setattr(RU, n, f)
        MSG.NPOINTS, MSG.RADIUS = saved_cfg
        U.set_kernels(None)
