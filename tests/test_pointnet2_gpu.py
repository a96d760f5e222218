"""PointNet++ primitives (csrc/pointnet2.cu) on the GPU.

Bar: IDENTICAL results (indices equal, floats bit-equal) to
  * the reference's own CUDA kernels, compiled from /root/reference/pycontrast/networks/pointnet2/src by oracle/build_ref.py into
    oracle/_ref/libpn2_ref.so (skipped when that library did not travel), and
  * the numpy restatement oracle/pn2_oracle.py (which the same run thereby pins to the reference kernels),
on clouds that contain duplicate points (the depth branch samples its 4096 points WITH replacement, build_backbone.py:421), so that
the tie rules matter.  The gradient kernels scatter with fp32 atomics in both implementations; they are compared on integer-valued
gradients, whose sums are exact in any order."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import pn2_oracle as PO

pytestmark = pytest.mark.gpu
DEV = "cuda"
REF_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libpn2_ref.so")


@pytest.fixture(scope="module")
def K():
    from hcmoco_b200.kernels import CudaKernels
    return CudaKernels()


@pytest.fixture(scope="module")
def REF():
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libpn2_ref.so not built (needs /root/reference at build time)")
    return ctypes.CDLL(REF_LIB)


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def cloud(B, N, seed, dup=0.3):
    """points in [-1,1]^3 on a coarse lattice (exact ties between distinct points) with a share of exact duplicates"""
    g = torch.Generator().manual_seed(seed)
    x = (torch.randint(-64, 65, (B, N, 3), generator=g).float() / 64.0)
    if dup > 0 and N > 4:
        src = torch.randint(0, N, (B, N), generator=g)
        take = torch.rand(B, N, generator=g) < dup
        x = torch.where(take[..., None], torch.gather(x, 1, src[..., None].expand(B, N, 3)), x)
    return x.contiguous().to(DEV)


@pytest.mark.parametrize("N,M", [(64, 64), (700, 256), (1024, 512), (4096, 1024), (5000, 300), (16384, 128), (3, 3)])
def test_fps_identical_to_reference_kernel(K, REF, N, M):
    B = 3
    xyz = cloud(B, N, N + M)
    mine = torch.zeros(B, M, dtype=torch.int32, device=DEV)
    K.pn2_furthest_point_sampling(xyz, B, N, M, mine)
    ref = torch.zeros(B, M, dtype=torch.int32, device=DEV)
    temp = torch.full((B, N), 1e10, device=DEV)
    REF.ref_furthest_point_sampling(B, N, M, P(xyz), P(temp), P(ref))
    torch.cuda.synchronize()
    assert torch.equal(mine, ref), (mine != ref).nonzero()[:5]
    if N <= 1024:
        assert np.array_equal(PO.furthest_point_sampling(xyz.cpu().numpy(), M), ref.cpu().numpy())


@pytest.mark.parametrize("N,M,radius,nsample", [(4096, 1024, 0.125, 16), (4096, 1024, 0.25, 32), (700, 300, 0.05, 16), (1024, 64, 1.5, 32),
                                                (33, 7, 0.3, 64)])
def test_ball_query_identical(K, REF, N, M, radius, nsample):
    B = 2
    xyz = cloud(B, N, 1)
    new_xyz = torch.cat([xyz[:, :M - 1], torch.full((B, 1, 3), 9.0, device=DEV)], 1).contiguous()      # the last query has no neighbour
    mine = torch.full((B, M, nsample), -7, dtype=torch.int32, device=DEV)
    K.pn2_ball_query(new_xyz, xyz, B, N, M, radius, nsample, mine)
    ref = torch.zeros(B, M, nsample, dtype=torch.int32, device=DEV)
    REF.ref_ball_query(B, N, M, ctypes.c_float(radius), nsample, P(new_xyz), P(xyz), P(ref))
    torch.cuda.synchronize()
    assert torch.equal(mine, ref)
    assert int(mine[:, -1].abs().sum()) == 0
    if N * M <= 1024 * 300:
        assert np.array_equal(PO.ball_query(radius, nsample, xyz.cpu().numpy(), new_xyz.cpu().numpy()), ref.cpu().numpy())


@pytest.mark.parametrize("n,m", [(65536, 4096), (1000, 700), (257, 2), (5, 1)])
def test_three_nn_and_interpolate_identical(K, REF, n, m):
    B, C = 2, 5
    known, unknown = cloud(B, m, 2), cloud(B, n, 3, dup=0.0)
    d1, i1 = torch.empty(B, n, 3, device=DEV), torch.zeros(B, n, 3, dtype=torch.int32, device=DEV)
    d2, i2 = torch.empty(B, n, 3, device=DEV), torch.zeros(B, n, 3, dtype=torch.int32, device=DEV)
    K.pn2_three_nn(unknown, known, B, n, m, d1, i1)
    REF.ref_three_nn(B, n, m, P(unknown), P(known), P(d2), P(i2))
    torch.cuda.synchronize()
    assert torch.equal(i1, i2) and torch.equal(d1, d2)
    if n <= 1000:
        od, oi = PO.three_nn(unknown.cpu().numpy(), known.cpu().numpy())
        assert np.array_equal(oi[:, :, :min(3, m)], i2.cpu().numpy()[:, :, :min(3, m)])
        assert np.array_equal(od[:, :, :min(3, m)], d2.cpu().numpy()[:, :, :min(3, m)])
    if m < 3:
        return
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(B, C, m, generator=g).to(DEV)
    w = torch.rand(B, n, 3, generator=g).to(DEV)
    o1, o2 = torch.empty(B, C, n, device=DEV), torch.empty(B, C, n, device=DEV)
    K.pn2_three_interpolate(feats, i1, w, B, C, m, n, o1)
    REF.ref_three_interpolate(B, C, m, n, P(feats), P(i2), P(w), P(o2))
    torch.cuda.synchronize()
    assert torch.equal(o1, o2)
    if n <= 1000:
        assert np.array_equal(PO.three_interpolate(feats.cpu().numpy(), i2.cpu().numpy(), w.cpu().numpy()), o2.cpu().numpy())
    # gradient: integer-valued grad_out and weights in {0, 1, 2} -> exact sums
    go = torch.randint(-3, 4, (B, C, n), generator=g).float().to(DEV)
    wi = torch.randint(0, 3, (B, n, 3), generator=g).float().to(DEV)
    g1, g2 = torch.zeros(B, C, m, device=DEV), torch.zeros(B, C, m, device=DEV)
    K.pn2_three_interpolate_grad(go, i1, wi, B, C, n, m, g1)
    REF.ref_three_interpolate_grad(B, C, n, m, P(go), P(i2), P(wi), P(g2))
    torch.cuda.synchronize()
    assert torch.equal(g1, g2)


@pytest.mark.parametrize("N,npoint,nsample", [(4096, 1024, 32), (300, 17, 5)])
def test_group_and_gather_identical(K, REF, N, npoint, nsample):
    B, C = 2, 19
    g = torch.Generator().manual_seed(N)
    feats = torch.randn(B, C, N, generator=g).to(DEV)
    idx = torch.randint(0, N, (B, npoint, nsample), generator=g, dtype=torch.int32).to(DEV)
    o1, o2 = torch.empty(B, C, npoint, nsample, device=DEV), torch.empty(B, C, npoint, nsample, device=DEV)
    K.pn2_group_points(feats, idx, B, C, N, npoint, nsample, o1)
    REF.ref_group_points(B, C, N, npoint, nsample, P(feats), P(idx), P(o2))
    go = torch.randint(-3, 4, (B, C, npoint, nsample), generator=g).float().to(DEV)
    g1, g2 = torch.zeros(B, C, N, device=DEV), torch.zeros(B, C, N, device=DEV)
    K.pn2_group_points_grad(go, idx, B, C, N, npoint, nsample, g1)
    REF.ref_group_points_grad(B, C, N, npoint, nsample, P(go), P(idx), P(g2))
    idx1 = idx[:, :, 0].contiguous()
    p1, p2 = torch.empty(B, C, npoint, device=DEV), torch.empty(B, C, npoint, device=DEV)
    K.pn2_gather_points(feats, idx1, B, C, N, npoint, p1)
    REF.ref_gather_points(B, C, N, npoint, P(feats), P(idx1), P(p2))
    go1 = go[:, :, :, 0].contiguous()
    h1, h2 = torch.zeros(B, C, N, device=DEV), torch.zeros(B, C, N, device=DEV)
    K.pn2_gather_points_grad(go1, idx1, B, C, N, npoint, h1)
    REF.ref_gather_points_grad(B, C, N, npoint, P(go1), P(idx1), P(h2))
    torch.cuda.synchronize()
    assert torch.equal(o1, o2) and torch.equal(g1, g2) and torch.equal(p1, p2) and torch.equal(h1, h2)
    assert np.array_equal(PO.group_points(feats.cpu().numpy(), idx.cpu().numpy()), o2.cpu().numpy())
    assert np.array_equal(PO.scatter_add(go.reshape(B, C, -1).cpu().numpy(), idx.reshape(B, -1).cpu().numpy(), N).astype(np.float32),
                          g2.cpu().numpy())


@pytest.mark.parametrize("lattice", [True, False])
def test_against_numpy_oracle_without_reference_library(K, lattice):
    """Runs also where oracle/_ref did not travel: the kernels against the numpy restatement on small clouds (lattice coordinates
    with exact ties, and generic floats where every product rounds: the oracle spells out nvcc's fma contraction)."""
    B, N, M = 2, 500, 64
    xyz = cloud(B, N, 9) if lattice else torch.randn(B, N, 3, generator=torch.Generator().manual_seed(3)).to(DEV)
    idx = torch.zeros(B, M, dtype=torch.int32, device=DEV)
    K.pn2_furthest_point_sampling(xyz, B, N, M, idx)
    assert np.array_equal(PO.furthest_point_sampling(xyz.cpu().numpy(), M), idx.cpu().numpy())
    new_xyz = torch.gather(xyz, 1, idx.long()[..., None].expand(B, M, 3)).contiguous()
    bq = torch.zeros(B, M, 16, dtype=torch.int32, device=DEV)
    K.pn2_ball_query(new_xyz, xyz, B, N, M, 0.3, 16, bq)
    assert np.array_equal(PO.ball_query(0.3, 16, xyz.cpu().numpy(), new_xyz.cpu().numpy()), bq.cpu().numpy())
    d, i = torch.empty(B, N, 3, device=DEV), torch.zeros(B, N, 3, dtype=torch.int32, device=DEV)
    K.pn2_three_nn(xyz, new_xyz, B, N, M, d, i)
    od, oi = PO.three_nn(xyz.cpu().numpy(), new_xyz.cpu().numpy())
    assert np.array_equal(oi, i.cpu().numpy()) and np.array_equal(od, d.cpu().numpy())


def test_fps_speed_vs_reference_kernel(K, REF):
    """Not a parity test: records the time of the Pointnet2MSG-sized call (B=32 clouds of 4096 points, 4096 / 1024 samples)."""
    B, N = 32, 4096
    xyz = cloud(B, N, 4)
    for M in (4096, 1024):
        mine = torch.zeros(B, M, dtype=torch.int32, device=DEV)
        ref = torch.zeros(B, M, dtype=torch.int32, device=DEV)
        temp = torch.full((B, N), 1e10, device=DEV)
        ts = []
        for fn in (lambda: K.pn2_furthest_point_sampling(xyz, B, N, M, mine),
                   lambda: (temp.fill_(1e10), REF.ref_furthest_point_sampling(B, N, M, P(xyz), P(temp), P(ref)))):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 3)
        print("fps B=%d N=%d M=%d: hcm %.3f ms, reference kernel %.3f ms (%.1fx)" % (B, N, M, ts[0], ts[1], ts[1] / ts[0]))
        assert torch.equal(mine, ref)
