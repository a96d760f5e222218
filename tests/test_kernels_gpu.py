"""Per-kernel numerics: every C-ABI entry point (CUDA, through ctypes) against the plain-PyTorch fp32
statement of the same op (tests/kernel_ref.py) on the same random inputs.  Tolerance 2e-5 relative
(fp32 sums in a different order) unless noted."""
import pytest
import torch

from kernel_ref import TorchKernels

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 2e-5


@pytest.fixture(scope="module")
def KK():
    from hcmoco_b200.kernels import CudaKernels
    # the PyTorch reference must be true fp32: cuDNN / cuBLAS default to TF32 on this GPU
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return CudaKernels(), TorchKernels(DEV)


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return (scale * torch.randn(*shape, generator=g)).to(DEV)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp(min=1e-20))


def both(KK, name, args, outs, tol=TOL):
    """Run kernel `name` with the CUDA backend and the reference on cloned arguments; compare args[outs]."""
    kc, kr = KK

    def cl(a):
        if not torch.is_tensor(a):
            return a
        if a.is_contiguous():
            return a.clone()
        c = torch.empty_strided(a.size(), a.stride(), dtype=a.dtype, device=a.device)   # keep row strides (ldx, ldo)
        c.copy_(a)
        return c

    a1 = [cl(a) for a in args]
    a2 = [cl(a) for a in args]
    getattr(kc, name)(*a1)
    getattr(kr, name)(*a2)
    torch.cuda.synchronize()
    for i in outs:
        e = rel(a1[i], a2[i])
        assert e < tol, (name, i, e)
    return a1, a2


CONVS = [  # B, H, W, Cin, Cout, ks, stride
    (2, 32, 32, 3, 64, 3, 2), (2, 16, 16, 64, 64, 3, 2), (2, 16, 16, 64, 256, 1, 1), (3, 16, 16, 256, 18, 3, 1),
    (2, 16, 16, 18, 18, 3, 1), (2, 8, 8, 36, 72, 3, 2), (5, 4, 4, 144, 144, 3, 1), (2, 6, 10, 32, 64, 1, 1),
    (1, 8, 8, 256, 256, 3, 1), (2, 16, 16, 72, 18, 1, 1),
]


def _out_hw(H, W, ks, s):
    p = (ks - 1) // 2
    return (H + 2 * p - ks) // s + 1, (W + 2 * p - ks) // s + 1


@pytest.mark.parametrize("shape", CONVS)
@pytest.mark.parametrize("onload", [False, True])
def test_conv_fwd(KK, shape, onload):
    B, H, W, Cin, Cout, ks, s = shape
    Ho, Wo = _out_hw(H, W, ks, s)
    x, w = rnd(B, H, W, Cin), rnd(Cout, Cin, ks, ks, scale=0.1)
    sc, sh = (rnd(Cin, seed=1).abs() + 0.5, rnd(Cin, seed=2)) if onload else (None, None)
    kc, kr = KK
    rows = kc.conv2d_stat_rows(B, H, W, Cin, Cout, ks, s)
    y1, y2 = torch.zeros(B, Ho, Wo, Cout, device=DEV), torch.zeros(B, Ho, Wo, Cout, device=DEV)
    p1, p2 = torch.zeros(rows, 2, Cout, device=DEV), torch.zeros(3, 2, Cout, device=DEV)
    kc.conv2d_fwd(x, w, None, y1, B, H, W, Cin, Cout, ks, s, sc, sh, 1, p1)
    kr.conv2d_fwd(x, w, None, y2, B, H, W, Cin, Cout, ks, s, sc, sh, 1, p2)
    assert rel(y1, y2) < TOL
    assert rel(p1.sum(0)[1], p2.sum(0)[1]) < TOL
    assert float((p1.sum(0)[0] - p2.sum(0)[0]).abs().max()) < 1e-4 * float(y2.abs().sum(dim=(0, 1, 2)).max())
    # bias, no stats
    bias = rnd(Cout, seed=5)
    kc.conv2d_fwd(x, w, bias, y1, B, H, W, Cin, Cout, ks, s, None, None, 0, None)
    kr.conv2d_fwd(x, w, bias, y2, B, H, W, Cin, Cout, ks, s, None, None, 0, None)
    assert rel(y1, y2) < TOL


@pytest.mark.parametrize("shape", CONVS)
def test_conv_dgrad_wgrad(KK, shape):
    B, H, W, Cin, Cout, ks, s = shape
    Ho, Wo = _out_hw(H, W, ks, s)
    x, w, dy = rnd(B, H, W, Cin), rnd(Cout, Cin, ks, ks, scale=0.1), rnd(B, Ho, Wo, Cout, seed=3)
    dx = rnd(B, H, W, Cin, seed=4)
    for acc in (0, 1):
        both(KK, "conv2d_dgrad", [dy, w, dx, B, H, W, Cin, Cout, ks, s, acc], [2])
    sc, sh = rnd(Cin, seed=1).abs() + 0.5, rnd(Cin, seed=2)
    dw = rnd(Cout, Cin, ks, ks, seed=6)
    both(KK, "conv2d_wgrad", [x, dy, dw, B, H, W, Cin, Cout, ks, s, None, None, 0], [2])
    both(KK, "conv2d_wgrad", [x, dy, dw, B, H, W, Cin, Cout, ks, s, sc, sh, 1], [2])


def test_gemm(KK):
    A, Bm, C = rnd(4 * 50 * 70), rnd(4 * 70 * 33, seed=1), rnd(4 * 50 * 33, seed=2)
    bias = rnd(33, seed=3)
    # plain batched row-major
    both(KK, "gemm", [A, Bm, None, C, 4, 50, 33, 70, 70, 1, 33, 1, 33, 50 * 70, 70 * 33, 50 * 33, 1.0, 0], [3])
    # A transposed, B transposed, bias, alpha, accumulate
    both(KK, "gemm", [A, Bm, bias, C, 4, 50, 33, 70, 1, 50, 1, 70, 33, 50 * 70, 70 * 33, 50 * 33, 0.5, 1], [3])
    # single tall-K product (the SemGCN / head weight gradients)
    A2, B2, C2 = rnd(300 * 256), rnd(300 * 128, seed=1), rnd(256 * 128, seed=2)
    both(KK, "gemm", [A2, B2, None, C2, 1, 256, 128, 300, 1, 256, 128, 1, 128, 0, 0, 0, 1.0, 0], [3])


@pytest.mark.parametrize("PC", [(2 * 16 * 16, 18), (3 * 5 * 7, 36), (4096, 64), (39, 128), (2 * 4 * 4, 144), (777, 256)])
def test_batchnorm(KK, PC):
    P, C = PC
    kc, kr = KK
    y = rnd(P, C) * 2 + 0.3
    gamma, beta = rnd(C, seed=1).abs() + 0.5, rnd(C, seed=2)
    outs = []
    for k in (kc, kr):
        rows = k.colstat_rows(P, C)
        part = torch.zeros(rows, 2, C, device=DEV)
        rm, rv, nbt = torch.zeros(C, device=DEV), torch.ones(C, device=DEV), torch.zeros((), dtype=torch.int64, device=DEV)
        sc, sh, mu, iv = (torch.zeros(C, device=DEV) for _ in range(4))
        k.bn_stats(y, P, C, part)
        k.bn_finalize(part, rows, C, P, gamma, beta, rm, rv, nbt, 0.01, 1e-5, sc, sh, mu, iv)
        res = rnd(P, C, seed=7)
        out = torch.zeros(P, C, device=DEV)
        k.bn_apply(y, sc, sh, res, gamma, beta, 1, out, P, C)
        # backward with the tensor mask and with the recomputed mask
        dz = rnd(P, C, seed=8)
        k1, k2, k3, dg, db = (torch.zeros(C, device=DEV) for _ in range(5))
        k.bn_bwd_reduce(dz, out, None, None, y, mu, iv, P, C, part)
        k.bn_bwd_finalize(part, rows, C, P, gamma, mu, iv, dg, db, k1, k2, k3)
        dy, go = torch.zeros(P, C, device=DEV), rnd(P, C, seed=9)
        k.bn_bwd_apply(dz, out, None, None, y, k1, k2, k3, dy, go, 1, P, C)
        k.bn_bwd_reduce(dz, None, sc, sh, y, mu, iv, P, C, part)
        k1b, k2b, k3b, dgb, dbb = (torch.zeros(C, device=DEV) for _ in range(5))
        k.bn_bwd_finalize(part, rows, C, P, gamma, mu, iv, dgb, dbb, k1b, k2b, k3b)
        dz2 = dz.clone()
        k.bn_bwd_apply(dz2, None, sc, sh, y, k1b, k2b, k3b, dz2, None, 0, P, C)     # in place
        outs.append((sc, sh, mu, iv, rm, rv, nbt.float(), out, dg, db, dy, go, dgb, dbb, dz2))
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(*outs)):
        assert rel(a, b) < 5e-5, (i, rel(a, b))


@pytest.mark.parametrize("PC", [(2 * 16 * 16, 18), (64 * 32 * 32, 36), (4096, 64), (39, 128), (8 * 16 * 16, 144), (777, 256),
                                (64 * 64 * 64, 18)])
def test_batchnorm_fused_finalize(KK, PC):
    """bn_stats_finalize / bn_bwd_reduce_finalize (the last CTA reduces the partial rows) == the two-launch forms, bit for bit,
    and the ticket counter is left re-armed (launched three times in a row)."""
    P, C = PC
    kc, _ = KK
    y = rnd(P, C) * 2 + 0.3
    gamma, beta = rnd(C, seed=1).abs() + 0.5, rnd(C, seed=2)
    dz = rnd(P, C, seed=8)
    rows = kc.colstat_rows(P, C)
    cnt = torch.zeros(4, dtype=torch.int32, device=DEV)

    def run(fused):
        part = torch.zeros(rows, 2, C, device=DEV)
        rm, rv, nbt = torch.zeros(C, device=DEV), torch.ones(C, device=DEV), torch.zeros((), dtype=torch.int64, device=DEV)
        sc, sh, mu, iv, k1, k2, k3, dg, db = (torch.zeros(C, device=DEV) for _ in range(9))
        for _ in range(3):
            if fused:
                kc.bn_stats_finalize(y, P, C, part, cnt, gamma, beta, rm, rv, nbt, 0.01, 1e-5, sc, sh, mu, iv)
                kc.bn_bwd_reduce_finalize(dz, None, sc, sh, y, mu, iv, P, C, part, cnt, gamma, dg, db, k1, k2, k3)
            else:
                kc.bn_stats(y, P, C, part)
                kc.bn_finalize(part, rows, C, P, gamma, beta, rm, rv, nbt, 0.01, 1e-5, sc, sh, mu, iv)
                kc.bn_bwd_reduce(dz, None, sc, sh, y, mu, iv, P, C, part)
                kc.bn_bwd_finalize(part, rows, C, P, gamma, mu, iv, dg, db, k1, k2, k3)
        torch.cuda.synchronize()
        return sc, sh, mu, iv, rm, rv, nbt.float(), k1, k2, k3, dg, db

    a, b = run(True), run(False)
    assert int(cnt[0]) == 0
    for i, (u, v) in enumerate(zip(a, b)):
        assert rel(u, v) < 1e-6, (i, rel(u, v))


def test_elementwise_and_layout(KK):
    a, b, g = rnd(1000), rnd(1000, seed=1), rnd(1000, seed=2)
    both(KK, "relu_bwd", [a, b, g, 0, 1000], [2])
    both(KK, "relu_bwd", [a, b, g, 1, 1000], [2])
    both(KK, "relu_bwd", [a, b, a, 0, 1000], [2])          # in place
    both(KK, "axpy", [a, b, 0.25, 1000], [0])
    x = rnd(3, 6, 8, 8)
    both(KK, "nchw_to_nhwc", [x, torch.zeros(3, 64, 3, device=DEV), 3, 6, 64, 3, 3], [1])
    f = rnd(3, 5, 7, 36)
    both(KK, "avgpool", [f, torch.zeros(3, 100, device=DEV), 3, 35, 36, 100, 18], [1])
    both(KK, "avgpool_bwd", [rnd(3, 100), rnd(3, 5, 7, 36, seed=1), 1, 3, 35, 36, 100, 18], [1])
    both(KK, "avgpool_bwd", [rnd(3, 100), rnd(3, 5, 7, 36, seed=1), 0, 3, 35, 36, 100, 18], [1])
    both(KK, "zero", [rnd(100), 200], [0])
    both(KK, "colsum_small", [rnd(7, 128), 7, 128, 128, rnd(128, seed=3), 1], [4])


@pytest.mark.parametrize("C", [18, 128])
def test_fuse_and_adjoint(KK, C):
    B, H, W = 2, 16, 16
    terms = [rnd(B, H, W, C), rnd(B, H // 2, W // 2, C, seed=1), rnd(B, H // 4, W // 4, C, seed=2),
             rnd(B, H // 8, W // 8, C, seed=3)]
    scales = [None, rnd(C, seed=4), rnd(C, seed=5), None]
    shifts = [None, rnd(C, seed=6), rnd(C, seed=7), None]
    log2f = torch.tensor([0, 1, 2, 3], dtype=torch.int32)
    out = torch.zeros(B, H, W, C, device=DEV)
    both(KK, "fuse_sum", [4, terms, scales, shifts, log2f, None, 1, out, B, H, W, C], [7])
    both(KK, "fuse_sum", [4, terms, None, None, log2f, rnd(C, seed=8), 0, out, B, H, W, C], [7])
    both(KK, "fuse_sum", [2, terms[:2], scales[:2], shifts[:2], log2f[:2].clone(), None, 1, out, B, H, W, C], [7])
    g = rnd(B, H, W, C, seed=9)
    for k in (1, 2, 3):
        o = rnd(B, H >> k, W >> k, C, seed=10)
        both(KK, "upsample_adjoint", [g, o, 0, B, H, W, C, k], [1])
        both(KK, "upsample_adjoint", [g, o, 1, B, H, W, C, k], [1])


def _banks(n):
    return [torch.nn.functional.normalize(rnd(n, 128, seed=s)) for s in (1, 2, 3)]


@pytest.mark.parametrize("BK", [(3, 257), (4, 1025), (2, 16385)])
def test_nce(KK, BK):
    B, K1 = BK
    n = 5000
    banks = _banks(n)
    f = torch.nn.functional.normalize(rnd(B, 3, 128), dim=2).reshape(B, 384)
    idx = torch.randint(0, n, (B, K1), generator=torch.Generator().manual_seed(0)).to(DEV)
    x = [f[:, 128 * m:128 * (m + 1)] for m in range(3)]
    kc, kr = KK
    res = []
    for use_depth in (None, torch.tensor([1, 0, 1, 1][:B], device=DEV), torch.zeros(B, dtype=torch.int64, device=DEV)):
        for k in (kc, kr):
            logits = torch.zeros(6, B, K1, device=DEV)
            lse, l0, hit, coef = (torch.zeros(6, B, device=DEV) for _ in range(4))
            loss, acc = torch.zeros(6, device=DEV), torch.zeros(6, device=DEV)
            df = torch.zeros(B, 384, device=DEV)
            k.nce_logits(*banks, *x, 384, idx, B, K1, 128, 0.07, logits)
            k.nce_loss(logits, B, K1, use_depth, None, lse, l0, hit, coef, loss, acc)
            k.nce_bwd(*banks, *x, 384, idx, B, K1, 128, 0.07, logits, lse, coef, 1.0, df, 384)
            res.append((logits, lse, loss, acc, coef, df))
        torch.cuda.synchronize()
        for i, (a, b) in enumerate(zip(res[-2], res[-1])):
            assert rel(a, b) < 1e-4 or float((a - b).abs().max()) < 1e-6, (i, rel(a, b))


def test_bank_update(KK):
    n = 1000
    bank = _banks(n)[0]
    x = torch.nn.functional.normalize(rnd(6, 3, 128), dim=2).reshape(6, 384)
    y = torch.tensor([5, 17, 5, 999, 0, 17], device=DEV)          # duplicates: the last writer wins
    both(KK, "bank_update", [bank, x[:, 128:256], 384, y, 6, 128, 0.5], [0], tol=1e-6)


def test_gather_l2norm(KK):
    B, HW, S = 3, 64, 50
    src = rnd(B, HW, 128)
    pix = torch.randint(0, HW, (B, S), generator=torch.Generator().manual_seed(0)).to(DEV)
    out, inv = torch.zeros(B * S, 128, device=DEV), torch.zeros(B * S, device=DEV)
    a1, _ = both(KK, "gather_l2norm", [src, 0, pix, HW, S, B * S, 128, out, 128, inv], [7, 9])
    dsrc = rnd(B, HW, 128, seed=3)
    both(KK, "gather_l2norm_bwd", [rnd(B * S, 128, seed=2), 128, a1[7], 128, a1[9], pix, HW, S, B * S, 128, dsrc, 0, 1], [10],
         tol=1e-4)
    # strided rows, no gather (the projection heads)
    lin = rnd(B, 128)
    f, inv2 = torch.zeros(B, 384, device=DEV), torch.zeros(B, device=DEV)
    a1, _ = both(KK, "gather_l2norm", [lin, 128, None, 0, 1, B, 128, f[:, 128:256], 384, inv2], [7, 9])
    dlin = torch.zeros(B, 128, device=DEV)
    both(KK, "gather_l2norm_bwd", [rnd(B, 384, seed=4)[:, 128:256], 384, a1[7], 384, a1[9], None, 0, 1, B, 128, dlin, 128, 0],
         [10], tol=1e-4)


@pytest.mark.parametrize("B,h,S,gs", [(4, 16, 60, (1.0, 1.0)), (3, 32, 400, (1.0, 1.0)), (2, 8, 130, (0.3, 2.0)), (5, 64, 400, (1.0, 0.0))])
def test_dense_affinity_fused(KK, B, h, S, gs):
    """Fused tcgen05 dense-affinity kernels (gather + L2-norm + S x S x 128 affinity + soft-target statistics; backward by
    recompute + second MMA + atomic scatter) against the unfused fp32 statement (contrast_trainer.py:684-723)."""
    kc, kr = KK
    g = torch.Generator().manual_seed(B + S)
    G1 = torch.randn(B, h * h, 128, generator=g).to(DEV)
    G2 = 0.5 * torch.randn(B, h * h, 128, generator=g).to(DEV) + 0.5 * G1
    pix = torch.randint(0, h * h, (B, S), generator=g).to(DEV)           # with replacement: duplicates as in the reference
    kept = (torch.rand(B, generator=g) < 0.7).float().to(DEV)
    kept[0] = 1
    use_depth = torch.ones(B, dtype=torch.int64, device=DEV)
    iT = 1.0 / 0.07
    out = []
    work = torch.empty((kc.dense_affinity_work_bytes(B, S) + 3) // 4, device=DEV)
    for kk in (kc, kr):
        stat, fin = torch.zeros(B, 2, S, 4, device=DEV), torch.zeros(8, device=DEV)
        kk.dense_affinity_fwd(G1, G2, pix, kept, use_depth, B, S, h, 128, iT, stat, fin, work)
        out.append((stat, fin))
    torch.cuda.synchronize()
    m = kept != 0
    for i in range(3):
        assert rel(out[0][0][m][..., i], out[1][0][m][..., i]) < 1e-5, i
    assert float((out[0][0][m][..., 3] != out[1][0][m][..., 3]).float().mean()) < 0.01      # argmax near-ties
    assert rel(out[0][1][:2], out[1][1][:2]) < 1e-5 and rel(out[0][1][2:5], out[1][1][2:5]) < 1e-2
    stat, fin = out[1]
    d = []
    for kk in (kc, kr):
        d1, d2 = torch.zeros_like(G1), torch.zeros_like(G2)
        kk.dense_affinity_bwd(G1, G2, pix, stat, kept, fin, B, S, h, 128, iT, gs[0], gs[1], d1, d2, work, int(kk is kc and S != 130))
        d.append((d1, d2))
    torch.cuda.synchronize()
    assert rel(d[0][0], d[1][0]) < 2e-4 and rel(d[0][1], d[1][1]) < 2e-4
    # all depth off -> exact zeros, and a backward that adds nothing
    stat, fin = torch.zeros(B, 2, S, 4, device=DEV), torch.ones(8, device=DEV)
    kc.dense_affinity_fwd(G1, G2, pix, kept, torch.zeros_like(use_depth), B, S, h, 128, iT, stat, fin, work)
    assert float(fin[:5].abs().sum()) == 0.0
    d1, d2 = torch.zeros_like(G1), torch.zeros_like(G2)
    kc.dense_affinity_bwd(G1, G2, pix, stat, kept, fin, B, S, h, 128, iT, gs[0], gs[1], d1, d2, work, 0)
    assert float(d1.abs().sum() + d2.abs().sum()) == 0.0


def test_stage2_loss_kernels(KK):
    kc, kr = KK
    B, J, h, S, R = 4, 13, 16, 60, 64
    use_depth = torch.tensor([1, 0, 1, 1], device=DEV)
    joints = (rnd(B, J, 2) * 40 + 30)
    both(KK, "joint_pixel_index", [joints, B * J, h, torch.zeros(B, J, dtype=torch.int64, device=DEV)], [3])
    mask = torch.zeros(B, R, R, device=DEV)
    mask[0, 10:40, 20:30] = 1
    mask[2, 1, 1] = 1            # not on the nearest-resize grid -> dropped
    mask[3, 4, 8] = 1
    both(KK, "dense_kept", [mask, B, R, h, torch.zeros(B, device=DEV)], [4])
    kept = torch.tensor([1., 0., 0., 1.], device=DEV)
    pix = torch.randint(0, h * h, (B, S), generator=torch.Generator().manual_seed(1)).to(DEV)
    Lm = rnd(B, S, S) * 3
    stat, fin = torch.zeros(B, 2, S, 4, device=DEV), torch.zeros(8, device=DEV)
    a1, _ = both(KK, "dense_stats", [Lm, pix, kept, use_depth, B, S, h, stat, fin], [7, 8], tol=1e-4)
    both(KK, "dense_grad", [Lm, pix, a1[7], kept, a1[8], B, S, h, 1.0], [0], tol=1e-4)
    # all depth off -> zeros
    a1, _ = both(KK, "dense_stats", [Lm, pix, kept, torch.zeros_like(use_depth), B, S, h, stat, fin], [8], tol=1e-4)
    assert float(a1[8][:5].abs().sum()) == 0.0
    vis = (torch.rand(B, J, generator=torch.Generator().manual_seed(2)) < 0.8).int().to(DEV)
    Lr, Ld = rnd(B, J, J) * 3, rnd(B, J, J, seed=1) * 3
    rs, lse, fin = torch.zeros(B, 2, 3, device=DEV), torch.zeros(B, 2, J, device=DEV), torch.zeros(8, device=DEV)
    a1, _ = both(KK, "joint_stats", [Lr, Ld, vis, use_depth, B, J, rs, lse, fin], [6, 7, 8], tol=1e-4)
    both(KK, "joint_grad", [Lr, Ld, vis, use_depth, a1[7], a1[8], B, J, 1.0], [0, 1], tol=1e-4)
    N = 2 * B * J
    Z = rnd(N, N) * 3
    rowstat, fin = torch.zeros(N, 3, device=DEV), torch.zeros(4, device=DEV)
    for ur in (None, torch.tensor([1, 1, 0, 1], device=DEV)):
        a1, _ = both(KK, "scl_stats", [Z, B, J, ur, use_depth, rowstat, fin], [5, 6], tol=1e-4)
        both(KK, "scl_grad", [Z, B, J, ur, use_depth, a1[5], a1[6], 1.0], [0], tol=1e-4)


@pytest.mark.parametrize("skeleton", ["mpii", "coco_reduce"])
def test_sgcn_kernels(KK, skeleton):
    from hcmoco_b200.layout import graph_edges
    J, rows, cols = graph_edges(skeleton)
    nnz = len(rows)
    rows_t = torch.tensor(rows, dtype=torch.int32, device=DEV)
    cols_t = torch.tensor(cols, dtype=torch.int32, device=DEV)
    e = rnd(1, nnz) + 1
    A = torch.zeros(J, J, device=DEV)
    a1, _ = both(KK, "sgcn_adj", [e, rows_t, cols_t, nnz, J, A], [5])
    A = a1[5]
    both(KK, "sgcn_adj_bwd", [A, rnd(J, J, seed=1), rows_t, cols_t, nnz, J, rnd(1, nnz, seed=2), 1], [6])
    B, Cin = 3, 128
    x = rnd(B, J, Cin)
    both(KK, "sgcn_aggregate", [x, A, B, J, Cin, torch.zeros(B, J, 2 * Cin, device=DEV)], [5])
    dxa = rnd(B, J, 2 * Cin, seed=3)
    both(KK, "sgcn_aggregate_bwd", [dxa, x, A, B, J, Cin, rnd(B, J, Cin, seed=4), 1, torch.zeros(J, J, device=DEV)], [6, 8],
         tol=1e-4)
    both(KK, "sgcn_aggregate_bwd", [dxa, x, A, B, J, Cin, None, 0, torch.zeros(J, J, device=DEV)], [8], tol=1e-4)
    both(KK, "joint_mean", [x, B, J, Cin, torch.zeros(B, Cin, device=DEV)], [4])
    both(KK, "joint_mean_bwd", [rnd(B, Cin), B, J, Cin, rnd(B, J, Cin, seed=5), 1], [4])


def test_sgd(KK):
    n = 10007
    p, g, buf = rnd(n + 1)[:n], rnd(n + 1, seed=1)[:n], rnd(n + 1, seed=2)[:n]
    both(KK, "sgd_step", [p, g, buf, n, 0.03, 0.9, 1e-4, 1, 1.0], [0, 2], tol=1e-6)
    both(KK, "sgd_step", [p, g, buf, n, 0.03, 0.9, 1e-4, 0, 0.125], [0, 2], tol=1e-6)


def test_missing_library_is_loud(monkeypatch):
    """No fallback: without the shared library the CUDA backend refuses to construct."""
    from hcmoco_b200 import _lib
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libhcmoco_sm100.so")
    monkeypatch.setattr(_lib, "_lib", None)
    from hcmoco_b200.kernels import CudaKernels
    with pytest.raises(_lib.HcmError):
        CudaKernels()


TC_CONVS = [  # B, H, W, Cin, Cout, ks
    (2, 16, 16, 18, 18, 3), (2, 16, 16, 64, 64, 3), (3, 8, 8, 144, 144, 3), (2, 16, 16, 64, 256, 1), (2, 16, 16, 256, 64, 1),
    (2, 16, 16, 256, 18, 3), (2, 32, 32, 36, 36, 3), (4, 64, 64, 18, 18, 3), (2, 8, 8, 72, 18, 1), (1, 12, 12, 256, 256, 3),
    (2, 16, 16, 32, 128, 1), (3, 4, 4, 128, 128, 3), (2, 96, 96, 32, 32, 3),
    # supertile geometries (several consecutive tiles per staged fill, contiguous tile ranges per CTA): st = 2, 4, 3
    (16, 64, 64, 18, 18, 3), (40, 64, 64, 18, 18, 3), (13, 96, 96, 32, 32, 3),
]


@pytest.mark.parametrize("shape", TC_CONVS)
def test_tc_conv(KK, shape):
    """tcgen05 bf16-split convolution against exact fp32 (F.conv2d, TF32 off).  Bar 3e-5 relative: the dropped
    lo*lo term and the bf16 rounding of lo are ~2^-17 per product."""
    B, H, W, Cin, Cout, ks = shape
    kc, kr = KK
    assert kc.tc_conv_supported(B, H, W, Cin, Cout, ks, 1)
    x, w = rnd(B, H, W, Cin), rnd(Cout, Cin, ks, ks, scale=0.1)
    sc, sh = rnd(Cin, seed=1).abs() + 0.5, rnd(Cin, seed=2)
    bias = rnd(Cout, seed=5)
    wp = torch.zeros((kc.tc_conv_wpack_bytes(B, H, W, Cin, Cout, ks) + 3) // 4, device=DEV)
    kc.tc_conv_pack(w, 0, wp, B, H, W, Cin, Cout, ks, 4 * kc.tc_conv_rowcat_supported(Cout, ks, 1))
    y1, y2 = rnd(B, H, W, Cout, seed=7), rnd(B, H, W, Cout, seed=7)
    # forward, BN+ReLU applied on load
    kc.tc_conv(x, wp, None, y1, B, H, W, Cin, Cout, ks, 1, sc, sh, 1, 0)
    kr.conv2d_fwd(x, w, None, y2, B, H, W, Cin, Cout, ks, 1, sc, sh, 1, None)
    assert rel(y1, y2) < 3e-5, rel(y1, y2)
    # plain + bias + accumulate
    kc.tc_conv(x, wp, bias, y1, B, H, W, Cin, Cout, ks, 1, None, None, 0, 1)
    y3 = torch.zeros_like(y2)
    kr.conv2d_fwd(x, w, bias, y3, B, H, W, Cin, Cout, ks, 1, None, None, 0, None)
    assert rel(y1, y2 + y3) < 3e-5
    # data gradient through the transposed pack
    if kc.tc_conv_supported(B, H, W, Cout, Cin, ks, 1):
        dy = rnd(B, H, W, Cout, seed=9)
        wpt = torch.zeros((kc.tc_conv_wpack_bytes(B, H, W, Cout, Cin, ks) + 3) // 4, device=DEV)
        kc.tc_conv_pack(w, 0, wpt, B, H, W, Cout, Cin, ks, 1 + 4 * kc.tc_conv_rowcat_supported(Cin, ks, 1))
        dx1, dx2 = torch.zeros(B, H, W, Cin, device=DEV), torch.zeros(B, H, W, Cin, device=DEV)
        kc.tc_conv(dy, wpt, None, dx1, B, H, W, Cout, Cin, ks, 1, None, None, 0, 0)
        kr.conv2d_dgrad(dy, w, dx2, B, H, W, Cin, Cout, ks, 1, 0)
        assert rel(dx1, dx2) < 3e-5, rel(dx1, dx2)


def test_tc_conv_rowcat():
    """The experimental row-concatenated formulation of tc_conv (HCM_TC_ROWCAT=1; off by default, see tc_conv.cu) stays
    parity-green: the same tc_conv cases in a child process with the switch on (the library reads it once per process)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, HCM_TC_ROWCAT="1")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_kernels_gpu.py"), "-q", "-x", "-m", "gpu", "-k",
                        "test_tc_conv and not rowcat and not stride2"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]


@pytest.mark.parametrize("shape", TC_CONVS)
def test_tc_wgrad(KK, shape):
    """tcgen05 weight gradient (both operands MN-major, accumulators resident in TMEM across tiles)."""
    B, H, W, Cin, Cout, ks = shape
    kc, kr = KK
    assert kc.tc_wgrad_supported(B, H, W, Cin, Cout, ks, 1)
    x, dy = rnd(B, H, W, Cin), rnd(B, H, W, Cout, seed=3)
    sc, sh = rnd(Cin, seed=1).abs() + 0.5, rnd(Cin, seed=2)
    dw1 = rnd(Cout, Cin, ks, ks, seed=6)
    dw2 = dw1.clone()
    kc.tc_wgrad(x, dy, dw1, 0, B, H, W, Cin, Cout, ks, 1, sc, sh, 1)
    kr.conv2d_wgrad(x, dy, dw2, B, H, W, Cin, Cout, ks, 1, sc, sh, 1)
    assert rel(dw1, dw2) < 3e-5, rel(dw1, dw2)
    kc.tc_wgrad(x, dy, dw1, 0, B, H, W, Cin, Cout, ks, 1, None, None, 0)
    kr.conv2d_wgrad(x, dy, dw2, B, H, W, Cin, Cout, ks, 1, None, None, 0)
    assert rel(dw1, dw2) < 3e-5, rel(dw1, dw2)


TC_S2 = [(2, 32, 32, 18, 18), (2, 16, 16, 18, 36), (3, 8, 8, 36, 72), (2, 64, 64, 64, 64), (2, 16, 16, 72, 144), (2, 8, 8, 18, 144),
         (4, 64, 64, 18, 18), (2, 32, 32, 32, 64), (2, 32, 32, 256, 36), (2, 8, 8, 128, 256), (2, 16, 16, 64, 128),
         (64, 64, 64, 18, 36)]       # (the last one: supertiles in the stride-2 data gradient)


@pytest.mark.parametrize("shape", TC_S2)
def test_tc_conv_stride2(KK, shape):
    """3x3 stride-2 convolution on tensor cores: the halo is staged space-to-depth (4 parity planes)."""
    B, H, W, Cin, Cout = shape
    kc, kr = KK
    assert kc.tc_conv_supported(B, H, W, Cin, Cout, 3, 2)
    x, w = rnd(B, H, W, Cin), rnd(Cout, Cin, 3, 3, scale=0.1)
    sc, sh = rnd(Cin, seed=1).abs() + 0.5, rnd(Cin, seed=2)
    wp = torch.zeros((kc.tc_conv_wpack_bytes(B, H, W, Cin, Cout, 3) + 3) // 4, device=DEV)
    kc.tc_conv_pack(w, 0, wp, B, H, W, Cin, Cout, 3, 0)
    y1, y2 = rnd(B, H // 2, W // 2, Cout, seed=7), rnd(B, H // 2, W // 2, Cout, seed=7)
    kc.tc_conv(x, wp, None, y1, B, H, W, Cin, Cout, 3, 2, sc, sh, 1, 0)
    kr.conv2d_fwd(x, w, None, y2, B, H, W, Cin, Cout, 3, 2, sc, sh, 1, None)
    assert rel(y1, y2) < 3e-5, rel(y1, y2)
    kc.tc_conv(x, wp, None, y1, B, H, W, Cin, Cout, 3, 2, None, None, 0, 0)
    kr.conv2d_fwd(x, w, None, y2, B, H, W, Cin, Cout, 3, 2, None, None, 0, None)
    assert rel(y1, y2) < 3e-5, rel(y1, y2)


@pytest.mark.parametrize("shape", TC_S2)
def test_tc_dgrad_stride2(KK, shape):
    """data gradient of the stride-2 conv: one 2x2-tap GEMM over dy, N = nqs output parities x Cin
    (4/nqs launches when 4*Cin does not fit N <= 256)."""
    B, H, W, Cin, Cout = shape
    kc, kr = KK
    if not kc.tc_dgrad_s2_supported(B, H, W, Cin, Cout):
        pytest.skip("ceil16(Cin) > 256")
    w, dy = rnd(Cout, Cin, 3, 3, scale=0.1), rnd(B, H // 2, W // 2, Cout, seed=3)
    wp = torch.zeros((kc.tc_dgrad_s2_wpack_bytes(B, H, W, Cin, Cout) + 3) // 4, device=DEV)
    kc.tc_dgrad_s2_pack(w, wp, B, H, W, Cin, Cout)
    dx1, dx2 = rnd(B, H, W, Cin, seed=4), rnd(B, H, W, Cin, seed=4)
    for acc in (0, 1):
        kc.tc_dgrad_s2(dy, wp, dx1, B, H, W, Cin, Cout, acc)
        kr.conv2d_dgrad(dy, w, dx2, B, H, W, Cin, Cout, 3, 2, acc)
        assert rel(dx1, dx2) < 3e-5, (acc, rel(dx1, dx2))


@pytest.mark.parametrize("shape", TC_S2)
def test_tc_wgrad_stride2(KK, shape):
    B, H, W, Cin, Cout = shape
    kc, kr = KK
    assert kc.tc_wgrad_supported(B, H, W, Cin, Cout, 3, 2)
    x, dy = rnd(B, H, W, Cin), rnd(B, H // 2, W // 2, Cout, seed=3)
    sc, sh = rnd(Cin, seed=1).abs() + 0.5, rnd(Cin, seed=2)
    dw1 = rnd(Cout, Cin, 3, 3, seed=6)
    dw2 = dw1.clone()
    kc.tc_wgrad(x, dy, dw1, 0, B, H, W, Cin, Cout, 3, 2, sc, sh, 1)
    kr.conv2d_wgrad(x, dy, dw2, B, H, W, Cin, Cout, 3, 2, sc, sh, 1)
    assert rel(dw1, dw2) < 3e-5, rel(dw1, dw2)


@pytest.mark.parametrize("R,Hs,Ws", [(256, 424, 512), (64, 100, 80)])
def test_stage_input(KK, R, Hs, Ws):
    """GPU input staging (datasets/dataset.py:104-160, 594-602) against its torch restatement: nearest-neighbour depth, mask and the
    millimetre sum / pixel count are bit-exact (integer work); the bilinear RGB planes and the mean-centred depth within fp32 rounding."""
    kc, kr = KK
    B = 5
    g = torch.Generator().manual_seed(R)
    rgb = torch.randint(0, 256, (B, Hs, Ws, 3), generator=g, dtype=torch.uint8).to(DEV)
    dmm = torch.randint(500, 4000, (B, Hs, Ws), generator=g, dtype=torch.int32)
    yy, xx = torch.meshgrid(torch.arange(Hs), torch.arange(Ws), indexing="ij")
    body = (((yy - Hs / 2) / (0.4 * Hs)) ** 2 + ((xx - Ws / 2) / (0.25 * Ws)) ** 2) <= 1.0
    dmm = (dmm * body).to(torch.uint16)
    dmm[3] = 0                                                     # an empty depth frame: count 0, mean 0, mask 0
    depth = dmm.to(DEV)
    # crop windows: inside the frame, partly outside (negative corner / past the border), whole frame
    crop = torch.tensor([[10, 20, Hs - 40, Ws - 60], [-15, -7, Hs // 2, Ws // 2], [Hs // 3, Ws // 3, Hs, Ws], [0, 0, Hs, Ws],
                         [5, 5, 37, 53]], dtype=torch.int32).to(DEV)
    flip = torch.tensor([0, 1, 0, 1, 1], dtype=torch.int32).to(DEV)
    has_depth = torch.tensor([1, 1, 1, 1, 0], dtype=torch.int64).to(DEV)
    outs = []
    for kk in (kc, kr):
        sums = torch.zeros(B, 2, dtype=torch.int64, device=DEV)
        x = torch.full((B, 6, R, R), float("nan"), device=DEV)
        mask = torch.full((B, R, R), float("nan"), device=DEV)
        kk.stage_input(rgb, depth, crop, flip, has_depth, B, Hs, Ws, R, sums, x, mask)
        outs.append((sums, x, mask))
    torch.cuda.synchronize()
    (s0, x0, m0), (s1, x1, m1) = outs
    assert torch.equal(s0, s1) and int(s1[0, 1]) > 0 and int(s1[3, 1]) == 0          # exact integer sums / counts
    assert torch.equal(m0, m1) and float(m1[4].sum()) == 0.0                         # mask exact; no depth -> empty mask
    assert torch.equal((x0[:, 3] != 0), (x1[:, 3] != 0))
    assert float((x0[:, 3:] - x1[:, 3:]).abs().max()) < 1e-6                         # same mm values, mean within fp32 rounding
    assert torch.equal(x0[:, 3], x0[:, 4]) and torch.equal(x0[:, 3], x0[:, 5])
    assert float((x0[:, :3] - x1[:, :3]).abs().max()) < 2e-5 and torch.isfinite(x0).all()
    # mean-centring: the masked depth of every depth-bearing sample sums to ~0
    for b in range(3):
        assert abs(float(x0[b, 3].double().sum())) / max(1.0, float(s1[b, 1])) < 1e-6


@pytest.mark.parametrize("P", [5, 4096, 70001])
def test_seg_head_kernels(KK, P):
    """seg_head.cu against its torch statement: the max of the two L2-normalised maps (forward, and backward routing), and the
    weighted / ignore-index cross entropy with the all-pixel top-1 accuracy."""
    kc, kr = KK
    m1, m2 = rnd(P, 128, seed=1), rnd(P, 128, seed=2)
    m1[0] = 0.0                                              # a zero row: the eps clamp of F.normalize
    for mb in (m2, None):
        outs = []
        for kk in (kc, kr):
            out, i1, i2 = torch.empty(P, 128, device=DEV), torch.empty(P, device=DEV), torch.empty(P, device=DEV)
            kk.l2norm_max_fwd(m1, mb, P, 128, out, i1, i2 if mb is not None else None)
            g = rnd(P, 128, seed=3)
            d1, d2 = rnd(P, 128, seed=4), rnd(P, 128, seed=5)
            kk.l2norm_max_bwd(g, m1, mb, i1, i2 if mb is not None else None, P, 128, 0.7, d1, d2 if mb is not None else None, 1)
            outs.append((out, i1[1:], d1[1:], d2[1:] if mb is not None else d1[1:]))
        torch.cuda.synchronize()
        for a, b in zip(*outs):
            assert rel(a, b) < TOL, rel(a, b)
    Cn = 25
    logits = rnd(P, Cn, seed=7, scale=3.0)
    g = torch.Generator().manual_seed(P)
    label = torch.randint(0, Cn, (P,), generator=g)
    label[torch.rand(P, generator=g) < 0.2] = 255
    label = label.to(DEV)
    cw = (torch.rand(Cn, generator=g) * 40 + 1).to(DEV)
    res = []
    for kk in (kc, kr):
        acc, out2, dl = torch.zeros(4, dtype=torch.float64, device=DEV), torch.empty(2, device=DEV), torch.empty(P, Cn, device=DEV)
        kk.seg_ce_fwd(logits, label, cw, P, Cn, 255, acc, out2)
        kk.seg_ce_bwd(logits, label, cw, P, Cn, 255, acc, 10.0, dl)
        res.append((acc[:3], out2, dl))
    torch.cuda.synchronize()
    assert rel(res[0][0], res[1][0]) < 1e-6 and float(res[0][0][2]) == float(res[1][0][2])       # hit count exact
    assert rel(res[0][1], res[1][1]) < 1e-6
    assert rel(res[0][2], res[1][2]) < TOL
    # all pixels ignored: loss 0, zero gradient (torch gives NaN)
    acc, out2, dl = torch.zeros(4, dtype=torch.float64, device=DEV), torch.empty(2, device=DEV), torch.empty(P, Cn, device=DEV)
    ign = torch.full((P,), 255, dtype=torch.int64, device=DEV)
    kc.seg_ce_fwd(logits, ign, cw, P, Cn, 255, acc, out2)
    kc.seg_ce_bwd(logits, ign, cw, P, Cn, 255, acc, 10.0, dl)
    assert float(out2[0]) == 0.0 and float(dl.abs().max()) == 0.0


def test_padded_stem_conv_on_tensor_cores(KK):
    """The 3-channel stem on the tensor-core kernels: input stored with a zero 4th channel (hcm_nchw_to_nhwc_pad), weights / weight
    gradient in the [64,3,3,3] checkpoint layout (pack and wgrad with ld = 3 < Cin = 4)."""
    kc, kr = KK
    B, R, Cout = 3, 64, 64
    x = rnd(B, 6, R, R)
    w = rnd(Cout, 3, 3, 3, scale=0.2)
    xin1, xin2 = torch.empty(B, R, R, 4, device=DEV), torch.empty(B, R, R, 4, device=DEV)
    kc.nchw_to_nhwc_pad(x, xin1, B, 6, R * R, 3, 3, 4)
    kr.nchw_to_nhwc_pad(x, xin2, B, 6, R * R, 3, 3, 4)
    assert torch.equal(xin1, xin2) and float(xin1[..., 3].abs().max()) == 0.0
    assert kc.tc_conv_supported(B, R, R, 4, Cout, 3, 2) and kc.tc_wgrad_supported(B, R, R, 4, Cout, 3, 2)
    wp = torch.zeros((kc.tc_conv_wpack_bytes(B, R, R, 4, Cout, 3) + 3) // 4, device=DEV)
    kc.tc_conv_pack(w, 3, wp, B, R, R, 4, Cout, 3, 0)
    y1 = torch.empty(B, R // 2, R // 2, Cout, device=DEV)
    kc.tc_conv(xin1, wp, None, y1, B, R, R, 4, Cout, 3, 2, None, None, 0, 0)
    y2 = torch.nn.functional.conv2d(x[:, 3:6], w, None, 2, 1).permute(0, 2, 3, 1)
    assert rel(y1, y2) < 3e-5, rel(y1, y2)
    dy = rnd(B, R // 2, R // 2, Cout, seed=3)
    dw1 = torch.zeros_like(w)
    guard = dw1.clone()
    kc.tc_wgrad(xin1, dy, dw1, 3, B, R, R, 4, Cout, 3, 2, None, None, 0)
    xr = x[:, 3:6].clone().requires_grad_(False)
    wr = w.clone().requires_grad_(True)
    (torch.nn.functional.conv2d(xr, wr, None, 2, 1) * dy.permute(0, 3, 1, 2)).sum().backward()
    assert rel(dw1, wr.grad) < 3e-5, rel(dw1, wr.grad)
    del guard
