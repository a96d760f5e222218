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


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        assert xyz.is_contiguous()
        K = _kernels()
        B, N, _ = xyz.size()
        out = _i32(K, B, npoint)
        K.pn2_furthest_point_sampling(xyz, B, N, npoint, out)
        return out

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        assert features.is_contiguous() and idx.is_contiguous()
        K = _kernels()
        B, npoint = idx.size()
        _, C, N = features.size()
        out = torch.empty(B, C, npoint, dtype=features.dtype, device=features.device)
        K.pn2_gather_points(features, idx, B, C, N, npoint, out)
        ctx.for_backwards = (idx, C, N)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        K = _kernels()
        B, npoint = idx.size()
        grad = torch.zeros(B, C, N, dtype=grad_out.dtype, device=grad_out.device)
        K.pn2_gather_points_grad(grad_out.contiguous(), idx, B, C, N, npoint, grad)
        return grad, None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown, known):
        assert unknown.is_contiguous() and known.is_contiguous()
        K = _kernels()
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = torch.empty(B, N, 3, dtype=unknown.dtype, device=unknown.device)
        idx = _i32(K, B, N, 3)
        K.pn2_three_nn(unknown, known, B, N, m, dist2, idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        assert features.is_contiguous() and idx.is_contiguous() and weight.is_contiguous()
        K = _kernels()
        B, c, m = features.size()
        n = idx.size(1)
        ctx.three_interpolate_for_backward = (idx, weight, m)
        out = torch.empty(B, c, n, dtype=features.dtype, device=features.device)
        K.pn2_three_interpolate(features, idx, weight, B, c, m, n, out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.three_interpolate_for_backward
        K = _kernels()
        B, c, n = grad_out.size()
        grad = torch.zeros(B, c, m, dtype=grad_out.dtype, device=grad_out.device)
        K.pn2_three_interpolate_grad(grad_out.contiguous(), idx, weight, B, c, n, m, grad)
        return grad, None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        assert features.is_contiguous() and idx.is_contiguous()
        K = _kernels()
        B, nfeatures, nsample = idx.size()
        _, C, N = features.size()
        out = torch.empty(B, C, nfeatures, nsample, dtype=features.dtype, device=features.device)
        K.pn2_group_points(features, idx, B, C, N, nfeatures, nsample, out)
        ctx.for_backwards = (idx, N)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, N = ctx.for_backwards
        K = _kernels()
        B, C, npoint, nsample = grad_out.size()
        grad = torch.zeros(B, C, N, dtype=grad_out.dtype, device=grad_out.device)
        K.pn2_group_points_grad(grad_out.contiguous(), idx, B, C, N, npoint, nsample, grad)
        return grad, None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        assert new_xyz.is_contiguous() and xyz.is_contiguous()
        K = _kernels()
        B, N, _ = xyz.size()
        npoint = new_xyz.size(1)
        idx = _i32(K, B, npoint, nsample)
        K.pn2_ball_query(new_xyz, xyz, B, N, npoint, radius, nsample, idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


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
