"""The three second-stage objectives with the reference's tensor-level signatures, DIFFERENTIABLE: each is a
torch.autograd.Function around the same forward / backward loss kernels the fused step uses, so that a reference-style
`loss = sum(losses); loss.backward()` back-propagates into the projection maps and the skeleton features exactly as
learning/contrast_trainer.py:642-723 (dense), :744-828 (sparse joint<->pixel), :830-892 (cross-subject SCL) do.
Inputs are the NCHW tensors the reference passes around; they are re-laid channels-last (a copy: plumbing) first."""
import torch


def _nhwc(K, t):
    return t.detach().to(K.device, K.dtype).permute(0, 2, 3, 1).contiguous()


def _i64(K, t, B, default=1):
    if t is None:
        return None
    return t.to(K.device).long().contiguous()


def _gscale(g):
    return 0.0 if g is None else float(g)


class _DenseFn(torch.autograd.Function):
    """contrast_trainer.py:642-723 through hcm_dense_affinity_{fwd,bwd} (gather + L2-norm + S x S affinity + soft-target
    log-softmax statistics fused; the backward recomputes the affinity)."""

    @staticmethod
    def forward(ctx, K, fm1, fm2, mask, ud, T, idx):
        B, C, h, _ = fm1.shape
        G1, G2 = _nhwc(K, fm1), _nhwc(K, fm2)
        S = idx.shape[1]
        kept = K.empty(B)
        K.dense_kept(mask, B, mask.shape[-1], h, kept)
        stat, fin = K.zeros(B, 2, S, 4), K.zeros(8)
        work = K.empty((K.dense_affinity_work_bytes(B, S) + 3) // 4)
        K.dense_affinity_fwd(G1, G2, idx, kept, ud, B, S, h, 128, 1.0 / T, stat, fin, work)
        ctx.K, ctx.T, ctx.dims = K, T, (B, S, h)
        ctx.save_for_backward(G1, G2, idx, stat, kept, fin, work)
        return fin[0].clone(), fin[1].clone(), fin[2].clone(), fin[3].clone()

    @staticmethod
    def backward(ctx, g0, g1, _a0, _a1):
        K, (B, S, h) = ctx.K, ctx.dims
        G1, G2, idx, stat, kept, fin, work = ctx.saved_tensors
        d1, d2 = torch.zeros_like(G1), torch.zeros_like(G2)
        K.dense_affinity_bwd(G1, G2, idx, stat, kept, fin, B, S, h, 128, 1.0 / ctx.T, _gscale(g0), _gscale(g1), d1, d2, work, 1)
        return None, d1.permute(0, 3, 1, 2), d2.permute(0, 3, 1, 2), None, None, None, None


def dense_loss(K, fm1, fm2, depth_mask, use_depth, T, S, sample_idx=None):
    B, C, h, w = fm1.shape
    assert C == 128 and h == w
    mask = depth_mask.detach().to(K.device, K.dtype).contiguous()
    R = mask.shape[-1]
    ud = _i64(K, use_depth, B)
    if sample_idx is None:
        step = R // h
        m = mask[:, ::step, ::step][:, :h, :h].reshape(B, -1)
        wts = torch.where(m.sum(1, keepdim=True) > 0, (m != 0).to(m.dtype), torch.ones_like(m))
        sample_idx = torch.multinomial(wts, S, replacement=True)
    idx = sample_idx.to(K.device).long().contiguous()
    l0, l1, a0, a1 = _DenseFn.apply(K, fm1, fm2, mask, ud, T, idx)
    return [l0, l1], [a0, a1]


def _joint_feats(K, G1, G2, joints_yx):
    """Rows [0,BJ): L2-normalised RGB-map features at the joints' pixels, rows [BJ,2BJ): the depth-map ones (+ what the
    backward needs: pixel ids, inverse norms)."""
    B, h = G1.shape[0], G1.shape[1]
    J = joints_yx.shape[1]
    pix = K.zeros(B, J, dtype=torch.int64)
    K.joint_pixel_index(joints_yx.detach().to(K.device, K.dtype).contiguous(), B * J, h, pix)
    Fm, inv = K.empty(2 * B * J, 128), K.empty(2 * B * J)
    K.gather_l2norm(G1, 0, pix, h * h, J, B * J, 128, Fm[:B * J], 128, inv[:B * J])
    K.gather_l2norm(G2, 0, pix, h * h, J, B * J, 128, Fm[B * J:], 128, inv[B * J:])
    return Fm, inv, pix, B, J, h


def _scatter_feats(K, dF, Fm, inv, pix, B, J, h, like1, like2):
    """Backward of _joint_feats: L2-norm backward + scatter-add of the joint rows into zeroed [B,h,h,128] maps."""
    d1, d2 = torch.zeros_like(like1), torch.zeros_like(like2)
    n = B * J
    K.gather_l2norm_bwd(dF[:n], 128, Fm[:n], 128, inv[:n], pix, h * h, J, n, 128, d1, 0, 1)
    K.gather_l2norm_bwd(dF[n:], 128, Fm[n:], 128, inv[n:], pix, h * h, J, n, 128, d2, 0, 1)
    return d1.permute(0, 3, 1, 2), d2.permute(0, 3, 1, 2)


class _JointFn(torch.autograd.Function):
    """contrast_trainer.py:744-828: logits[b][k][j] = <skeleton_k, pixel-feature_j>/T, CE over k with target j."""

    @staticmethod
    def forward(ctx, K, fm1, fm2, skel, joints_yx, vis, ud, T):
        G1, G2 = _nhwc(K, fm1), _nhwc(K, fm2)
        Fm, inv, pix, B, J, h = _joint_feats(K, G1, G2, joints_yx)
        n = B * J
        sk = skel.detach().to(K.device, K.dtype).contiguous()
        Sk, inv_s = K.empty(n, 128), K.empty(n)
        K.gather_l2norm(sk, 128, None, 0, 1, n, 128, Sk, 128, inv_s)
        Lr, Ld = K.empty(B, J, J), K.empty(B, J, J)
        K.gemm(Sk, Fm[:n], None, Lr, B, J, J, 128, 128, 1, 1, 128, J, J * 128, J * 128, J * J, 1.0 / T, 0)
        K.gemm(Sk, Fm[n:], None, Ld, B, J, J, 128, 128, 1, 1, 128, J, J * 128, J * 128, J * J, 1.0 / T, 0)
        rs, lse, fin = K.empty(B, 2, 3), K.empty(B, 2, J), K.zeros(8)
        K.joint_stats(Lr, Ld, vis, ud, B, J, rs, lse, fin)
        ctx.K, ctx.T, ctx.dims = K, T, (B, J, h)
        ctx.ud = ud
        ctx.save_for_backward(G1, G2, Fm, inv, pix, Sk, inv_s, Lr, Ld, lse, fin, vis)
        return fin[0].clone(), fin[1].clone(), fin[2].clone(), fin[3].clone()

    @staticmethod
    def backward(ctx, g0, g1, _a0, _a1):
        K, (B, J, h), iT = ctx.K, ctx.dims, 1.0 / ctx.T
        G1, G2, Fm, inv, pix, Sk, inv_s, Lr, Ld, lse, fin, vis = ctx.saved_tensors
        n = B * J
        dLr, dLd = Lr.clone(), Ld.clone()
        K.joint_grad(dLr, dLd, vis, ctx.ud, lse, fin, B, J, 1.0)          # logits -> d(loss_rgb)/dLr, d(loss_depth)/dLd in place
        dLr.mul_(_gscale(g0))
        dLd.mul_(_gscale(g1))
        dSk, dF = K.empty(n, 128), K.empty(2 * n, 128)
        Fa, Fd = Fm[:n], Fm[n:]
        # dSk[b][k] = iT * (sum_j dLr[k][j] a_j + sum_j dLd[k][j] d_j);  dFa[b][j] = iT * sum_k dLr[k][j] s_k
        K.gemm(dLr, Fa, None, dSk, B, J, 128, J, J, 1, 128, 1, 128, J * J, J * 128, J * 128, iT, 0)
        K.gemm(dLd, Fd, None, dSk, B, J, 128, J, J, 1, 128, 1, 128, J * J, J * 128, J * 128, iT, 1)
        K.gemm(dLr, Sk, None, dF[:n], B, J, 128, J, 1, J, 128, 1, 128, J * J, J * 128, J * 128, iT, 0)
        K.gemm(dLd, Sk, None, dF[n:], B, J, 128, J, 1, J, 128, 1, 128, J * J, J * 128, J * 128, iT, 0)
        d1, d2 = _scatter_feats(K, dF, Fm, inv, pix, B, J, h, G1, G2)
        dsk = K.empty(B, J, 128)
        K.gather_l2norm_bwd(dSk, 128, Sk, 128, inv_s, None, 0, 1, n, 128, dsk, 128, 0)
        return None, d1, d2, dsk, None, None, None, None


def joint_loss(K, fm1, fm2, skeleton_map, joints_yx, joints_vis, use_depth, T):
    B = fm1.shape[0]
    vis = joints_vis.to(K.device).int().contiguous()
    l0, l1, a0, a1 = _JointFn.apply(K, fm1, fm2, skeleton_map, joints_yx, vis, _i64(K, use_depth, B), T)
    return [l0, l1], [a0, a1]


class _SclFn(torch.autograd.Function):
    """contrast_trainer.py:830-892 (use_rgb None => all ones, segment_trainer.py:601-606)."""

    @staticmethod
    def forward(ctx, K, fm1, fm2, joints_yx, ur, ud, T):
        G1, G2 = _nhwc(K, fm1), _nhwc(K, fm2)
        Fm, inv, pix, B, J, h = _joint_feats(K, G1, G2, joints_yx)
        N = 2 * B * J
        Z, rowstat, fin = K.empty(N, N), K.empty(N, 3), K.zeros(4)
        K.gemm(Fm, Fm, None, Z, 1, N, N, 128, 128, 1, 1, 128, N, 0, 0, 0, 1.0 / T, 0)
        K.scl_stats(Z, B, J, ur, ud, rowstat, fin)
        ctx.K, ctx.T, ctx.dims, ctx.ur, ctx.ud = K, T, (B, J, h), ur, ud
        ctx.save_for_backward(G1, G2, Fm, inv, pix, Z, rowstat, fin)
        return fin[0].clone()

    @staticmethod
    def backward(ctx, g):
        K, (B, J, h), iT = ctx.K, ctx.dims, 1.0 / ctx.T
        G1, G2, Fm, inv, pix, Z, rowstat, fin = ctx.saved_tensors
        N = 2 * B * J
        dZ = Z.clone()
        K.scl_grad(dZ, B, J, ctx.ur, ctx.ud, rowstat, fin, _gscale(g))
        dF = K.zeros(N, 128)
        # dF = iT * (dZ + dZ^T) F
        K.gemm(dZ, Fm, None, dF, 1, N, 128, N, N, 1, 128, 1, 128, 0, 0, 0, iT, 1)
        K.gemm(dZ, Fm, None, dF, 1, N, 128, N, 1, N, 128, 1, 128, 0, 0, 0, iT, 1)
        d1, d2 = _scatter_feats(K, dF, Fm, inv, pix, B, J, h, G1, G2)
        return None, d1, d2, None, None, None, None


def scl_loss(K, fm1, fm2, joints_yx, use_depth, use_rgb, T):
    B = fm1.shape[0]
    return [_SclFn.apply(K, fm1, fm2, joints_yx, _i64(K, use_rgb, B), _i64(K, use_depth, B), T)], []
