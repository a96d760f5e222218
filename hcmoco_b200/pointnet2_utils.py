"""PointNet++ primitives with the reference's Python surface (pycontrast/networks/pointnet2/pointnet2_utils.py:10-290), backed by
the hcm_pn2_* kernels (csrc/pointnet2.cu) instead of the reference's `pointnet2_cuda` extension.

Same names, argument order, shapes, dtypes (int32 indices) and gradients as the reference:
    furthest_point_sample(xyz [B,N,3], npoint) -> [B,npoint] int32
    gather_operation(features [B,C,N], idx [B,npoint]) -> [B,C,npoint]                        (differentiable in features)
    three_nn(unknown [B,n,3], known [B,m,3]) -> (dist [B,n,3] = sqrt of the squared distances, idx [B,n,3] int32)
    three_interpolate(features [B,c,m], idx [B,n,3], weight [B,n,3]) -> [B,c,n]                (differentiable in features)
    grouping_operation(features [B,C,N], idx [B,npoint,nsample]) -> [B,C,npoint,nsample]       (differentiable in features)
    ball_query(radius, nsample, xyz [B,N,3], new_xyz [B,npoint,3]) -> [B,npoint,nsample] int32
    QueryAndGroup, GroupAll — the two grouper modules of pointnet2_modules.py
`networks/pointnet2_msg.py` / `pointnet2_modules.py` of the reference run on these unchanged (they are plain nn.Modules over the six
functions); the HRNetPN model wiring (networks/build_backbone.py:305-514) itself is not part of this package.
No CPU fallback: the kernels object is `CudaKernels` unless a test injects its reference executor (`set_kernels`)."""
import torch
import torch.nn as nn
from torch.autograd import Function

_K = None


def set_kernels(K):
    global _K
    _K = K


def _kernels():
    global _K
    if _K is None:
        from .kernels import CudaKernels
        _K = CudaKernels()
    return _K


def _i32(K, *shape):
    return torch.zeros(*shape, dtype=torch.int32, device=K.device)


# ---- index producers (no gradient): plain functions
@torch.no_grad()
def furthest_point_sample(xyz, npoint):
    """xyz [B,N,3] contiguous -> int32 [B,npoint]: iterative farthest point sampling from point 0 (pointnet2_utils.py:12-29)."""
    assert xyz.is_contiguous()
    K, (B, N, _) = _kernels(), xyz.shape
    out = _i32(K, B, npoint)
    K.pn2_furthest_point_sampling(xyz, B, N, npoint, out)
    return out


@torch.no_grad()
def three_nn(unknown, known):
    """unknown [B,n,3], known [B,m,3] -> (distances [B,n,3] (square roots), int32 indices [B,n,3]) (pointnet2_utils.py:79-99)."""
    assert unknown.is_contiguous() and known.is_contiguous()
    K, (B, n, _), m = _kernels(), unknown.shape, known.shape[1]
    d2, idx = torch.empty_like(unknown), _i32(K, B, n, 3)
    K.pn2_three_nn(unknown, known, B, n, m, d2, idx)
    return d2.sqrt_(), idx


@torch.no_grad()
def ball_query(radius, nsample, xyz, new_xyz):
    """int32 [B,npoint,nsample]: the first nsample points of xyz within `radius` of each new_xyz (pointnet2_utils.py:203-221)."""
    assert new_xyz.is_contiguous() and xyz.is_contiguous()
    K, (B, N, _), npoint = _kernels(), xyz.shape, new_xyz.shape[1]
    idx = _i32(K, B, npoint, nsample)
    K.pn2_ball_query(new_xyz, xyz, B, N, npoint, float(radius), int(nsample), idx)
    return idx


# ---- differentiable gathers: features [B,C,N] indexed along the point axis; the backward scatters into a zeroed buffer
class _Gather(Function):
    """gather_operation (idx [B,npoint]) and grouping_operation (idx [B,npoint,nsample]) are the same kernel family."""

    @staticmethod
    def forward(ctx, features, idx):
        assert features.is_contiguous() and idx.is_contiguous()
        K, (B, C, N) = _kernels(), features.shape
        out = features.new_empty((B, C) + tuple(idx.shape[1:]))
        if idx.dim() == 2:
            K.pn2_gather_points(features, idx, B, C, N, idx.shape[1], out)
        else:
            K.pn2_group_points(features, idx, B, C, N, idx.shape[1], idx.shape[2], out)
        ctx.idx, ctx.N = idx, N
        return out

    @staticmethod
    def backward(ctx, grad_out):
        K, idx = _kernels(), ctx.idx
        B, C = grad_out.shape[:2]
        grad = grad_out.new_zeros(B, C, ctx.N)
        g = grad_out.contiguous()
        if idx.dim() == 2:
            K.pn2_gather_points_grad(g, idx, B, C, ctx.N, idx.shape[1], grad)
        else:
            K.pn2_group_points_grad(g, idx, B, C, ctx.N, idx.shape[1], idx.shape[2], grad)
        return grad, None


def gather_operation(features, idx):
    return _Gather.apply(features, idx)


def grouping_operation(features, idx):
    return _Gather.apply(features, idx)


class _Interpolate(Function):
    """three_interpolate: out[b,c,p] = sum_j weight[b,p,j] * features[b,c,idx[b,p,j]] (pointnet2_utils.py:111-151)."""

    @staticmethod
    def forward(ctx, features, idx, weight):
        assert features.is_contiguous() and idx.is_contiguous() and weight.is_contiguous()
        K, (B, c, m), n = _kernels(), features.shape, idx.shape[1]
        out = features.new_empty(B, c, n)
        K.pn2_three_interpolate(features, idx, weight, B, c, m, n, out)
        ctx.saved = (idx, weight, m)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.saved
        K, (B, c, n) = _kernels(), grad_out.shape
        grad = grad_out.new_zeros(B, c, m)
        K.pn2_three_interpolate_grad(grad_out.contiguous(), idx, weight, B, c, n, m, grad)
        return grad, None, None


def three_interpolate(features, idx, weight):
    return _Interpolate.apply(features, idx, weight)


class QueryAndGroup(nn.Module):            # pointnet2_utils.py:231-264
    def __init__(self, radius, nsample, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        xyz_trans = xyz.transpose(1, 2).contiguous()
        grouped_xyz = grouping_operation(xyz_trans, idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is not None:
            grouped_features = grouping_operation(features, idx)
            return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
        return grouped_xyz


class GroupAll(nn.Module):                 # pointnet2_utils.py:267-290
    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        return grouped_xyz
