"""The three second-stage objectives with the reference's tensor-level signatures (forward values), running the
same loss kernels as the fused step: learning/contrast_trainer.py:642-723 (dense), :744-828 (sparse), :830-892 (SCL).
Inputs are the NCHW tensors the reference passes around; they are re-laid channels-last (a copy: plumbing) first."""
import torch


def _nhwc(K, t):
    return t.detach().to(K.device, K.dtype).permute(0, 2, 3, 1).contiguous()


def _i64(K, t, B, default=1):
    if t is None:
        return None
    return t.to(K.device).long().contiguous()


def dense_loss(K, fm1, fm2, depth_mask, use_depth, T, S, sample_idx=None):
    B, C, h, w = fm1.shape
    assert C == 128 and h == w
    G1, G2 = _nhwc(K, fm1), _nhwc(K, fm2)
    mask = depth_mask.detach().to(K.device, K.dtype).contiguous()
    R = mask.shape[-1]
    ud = _i64(K, use_depth, B)
    kept = K.empty(B)
    K.dense_kept(mask, B, R, h, kept)
    if sample_idx is None:
        step = R // h
        m = mask[:, ::step, ::step][:, :h, :h].reshape(B, -1)
        wts = torch.where(m.sum(1, keepdim=True) > 0, (m != 0).to(m.dtype), torch.ones_like(m))
        sample_idx = torch.multinomial(wts, S, replacement=True)
    idx = sample_idx.to(K.device).long().contiguous()
    S = idx.shape[1]
    A, D = K.empty(B * S, 128), K.empty(B * S, 128)
    K.gather_l2norm(G1, 0, idx, h * h, S, B * S, 128, A, 128, None)
    K.gather_l2norm(G2, 0, idx, h * h, S, B * S, 128, D, 128, None)
    Lm, stat, fin = K.empty(B, S, S), K.empty(B, 2, S, 4), K.zeros(8)
    K.gemm(D, A, None, Lm, B, S, S, 128, 128, 1, 1, 128, S, S * 128, S * 128, S * S, 1.0 / T, 0)
    K.dense_stats(Lm, idx, kept, ud, B, S, h, stat, fin)
    return [fin[0], fin[1]], [fin[2], fin[3]]


def _joint_feats(K, fm1, fm2, joints_yx):
    B, C, h, w = fm1.shape
    J = joints_yx.shape[1]
    G1, G2 = _nhwc(K, fm1), _nhwc(K, fm2)
    pix = K.zeros(B, J, dtype=torch.int64)
    K.joint_pixel_index(joints_yx.detach().to(K.device, K.dtype).contiguous(), B * J, h, pix)
    Fm = K.empty(2 * B * J, 128)
    K.gather_l2norm(G1, 0, pix, h * h, J, B * J, 128, Fm[:B * J], 128, None)
    K.gather_l2norm(G2, 0, pix, h * h, J, B * J, 128, Fm[B * J:], 128, None)
    return Fm, B, J


def joint_loss(K, fm1, fm2, skeleton_map, joints_yx, joints_vis, use_depth, T):
    Fm, B, J = _joint_feats(K, fm1, fm2, joints_yx)
    Sk = K.empty(B * J, 128)
    K.gather_l2norm(skeleton_map.detach().to(K.device, K.dtype).contiguous(), 128, None, 0, 1, B * J, 128, Sk, 128, None)
    Lr, Ld = K.empty(B, J, J), K.empty(B, J, J)
    K.gemm(Sk, Fm[:B * J], None, Lr, B, J, J, 128, 128, 1, 1, 128, J, J * 128, J * 128, J * J, 1.0 / T, 0)
    K.gemm(Sk, Fm[B * J:], None, Ld, B, J, J, 128, 128, 1, 1, 128, J, J * 128, J * 128, J * J, 1.0 / T, 0)
    rs, lse, fin = K.empty(B, 2, 3), K.empty(B, 2, J), K.zeros(8)
    K.joint_stats(Lr, Ld, joints_vis.to(K.device).int().contiguous(), _i64(K, use_depth, B), B, J, rs, lse, fin)
    return [fin[0], fin[1]], [fin[2], fin[3]]


def scl_loss(K, fm1, fm2, joints_yx, use_depth, use_rgb, T):
    Fm, B, J = _joint_feats(K, fm1, fm2, joints_yx)
    N = 2 * B * J
    Z, rowstat, fin = K.empty(N, N), K.empty(N, 3), K.zeros(4)
    K.gemm(Fm, Fm, None, Z, 1, N, N, 128, 128, 1, 1, 128, N, 0, 0, 0, 1.0 / T, 0)
    K.scl_stats(Z, B, J, _i64(K, use_rgb, B), _i64(K, use_depth, B), rowstat, fin)
    return [fin[0]], []
