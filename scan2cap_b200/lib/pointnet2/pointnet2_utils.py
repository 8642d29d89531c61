"""Mirror of the reference's lib/pointnet2/pointnet2_utils.py: the same autograd Functions
(FurthestPointSampling :51, GatherOperation :83, ThreeNN :120, ThreeInterpolate :152, GroupingOperation :209,
BallQuery :260), functional aliases, QueryAndGroup :294 and GroupAll :379 -- all backed by libs2c.so.

Addition: ``query_and_group`` / ``QueryAndGroup.forward`` run the FUSED kernel (ball query + both gathers +
centre subtraction + 1/radius + concat in one launch, s2c_query_and_group) whenever the reference's
sample_uniformly option is off, with a hand-written backward (one scatter-add for the features and, when xyz
requires grad -- the vote-aggregation layer -- one for the coordinates).
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _ext, _ext_mlp


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        fps_inds = _ext.furthest_point_sampling(xyz, npoint)
        ctx.mark_non_differentiable(fps_inds)
        return fps_inds

    @staticmethod
    def backward(xyz, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        _, C, N = features.size()
        ctx.for_backwards = (idx, C, N)
        return _ext.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        return _ext.gather_points_grad(grad_out.contiguous(), idx, N), None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown, known):
        dist2, idx = _ext.three_nn(unknown, known)
        ctx.mark_non_differentiable(dist2, idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        m = features.size(2)
        ctx.three_interpolate_for_backward = (idx, weight, m)
        return _ext.three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.three_interpolate_for_backward
        return _ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, m), None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        N = features.size(2)
        ctx.for_backwards = (idx, N)
        return _ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, N = ctx.for_backwards
        return _ext.group_points_grad(grad_out.contiguous(), idx, N), None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        inds = _ext.ball_query(new_xyz, xyz, radius, nsample)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class _FusedQueryAndGroup(Function):
    """grouped (B,3+C,M,ns) = cat([ (xyz[idx]-new_xyz) * (1/r if normalize), features[idx] ]) in one kernel."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, features, radius, nsample, normalize_xyz, feat_point_major, channels_last, pad4,
                grid=None):
        grouped, idx = _ext.query_and_group(xyz, new_xyz, features, radius, nsample, normalize_xyz,
                                            feat_point_major=feat_point_major, channels_last=channels_last, pad4=pad4,
                                            grid=grid)
        ctx.idx = idx
        ctx.n = xyz.shape[1]
        ctx.scale = (1.0 / radius) if normalize_xyz else 1.0
        ctx.feat_point_major = feat_point_major
        ctx.has_feat = features is not None
        ctx.feat_col = 4 if (pad4 and channels_last) else 3  # first feature column of the grouped rows
        ctx.C = 0 if features is None else (features.shape[2] if feat_point_major else features.shape[1])
        ctx.mark_non_differentiable(idx)
        return grouped, idx

    @staticmethod
    def backward(ctx, grad, _grad_idx):
        idx, n = ctx.idx, ctx.n
        g_xyz = g_new = g_feat = None
        B, Cp, M, ns = grad.shape
        rows = grad.permute(0, 2, 3, 1)  # channels-last storage (what the fused MLP backward hands back) -> (B,M,ns,Cp)
        if rows.is_contiguous() and grad.dtype == torch.float32:
            # one scatter-add launch per gathered tensor, straight from the channels-last rows into point-major
            # gradients: no (B,C,M,ns) re-layout copies, contiguous (vectorised) atomics per neighbour
            rows = rows.reshape(B, M * ns, Cp)
            if ctx.needs_input_grad[0]:
                g_xyz = _ext_mlp.group_rows_grad(rows, 0, 3, idx, n, ctx.scale)  # (B,n,3)
            if ctx.needs_input_grad[1]:
                g_new = -(grad[:, :3].sum(-1) * ctx.scale).transpose(1, 2)
            if ctx.has_feat and ctx.needs_input_grad[2]:
                g_feat = _ext_mlp.group_rows_grad(rows, ctx.feat_col, ctx.C, idx, n)  # (B,n,C)
                if not ctx.feat_point_major:
                    g_feat = g_feat.transpose(1, 2)
            return g_xyz, g_new, g_feat, None, None, None, None, None, None, None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            gx = grad[:, :3] * ctx.scale  # (B,3,M,ns)
            if ctx.needs_input_grad[0]:
                g_xyz = _ext.group_points_grad(gx.contiguous(), idx, n).transpose(1, 2)
            if ctx.needs_input_grad[1]:
                g_new = -gx.sum(-1).transpose(1, 2)
        if ctx.has_feat and ctx.needs_input_grad[2]:
            g_feat = _ext.group_points_grad(grad[:, ctx.feat_col:ctx.feat_col + ctx.C].contiguous(), idx, n)  # (B,C,n)
            if ctx.feat_point_major:
                g_feat = g_feat.transpose(1, 2)
        return g_xyz, g_new, g_feat, None, None, None, None, None, None, None


def query_and_group(xyz, new_xyz, features, radius, nsample, normalize_xyz=False, feat_point_major=False,
                    channels_last=False, pad4=False, grid=None):
    """grid: the uniform grid of (xyz, radius) built ahead of time (_ext.ball_query_grid_build), or None."""
    return _FusedQueryAndGroup.apply(xyz, new_xyz, features, radius, nsample, normalize_xyz, feat_point_major,
                                     channels_last, pad4, grid)


class QueryAndGroup(nn.Module):
    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False,
                 sample_uniformly=False, ret_unique_cnt=False):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz
        self.normalize_xyz = normalize_xyz
        self.sample_uniformly = sample_uniformly
        self.ret_unique_cnt = ret_unique_cnt
        if self.ret_unique_cnt:
            assert self.sample_uniformly
        if self.sample_uniformly:
            raise NotImplementedError("sample_uniformly is a host-side torch.unique/randint loop in the reference "
                                      "(pointnet2_utils.py:336-345) that CapNet never enables; out of scope")

    def forward(self, xyz, new_xyz, features=None):
        if not self.use_xyz:
            idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
            assert features is not None, "Cannot have not features and not use xyz as a feature!"
            new_features = grouping_operation(features, idx)
            if not self.ret_grouped_xyz:
                return new_features
            grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
            grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
            if self.normalize_xyz:
                grouped_xyz = grouped_xyz / self.radius
            return new_features, grouped_xyz
        new_features, _ = query_and_group(xyz, new_xyz, features, self.radius, self.nsample, self.normalize_xyz)
        if self.ret_grouped_xyz:
            return new_features, new_features[:, :3]
        return new_features


class GroupAll(nn.Module):
    def __init__(self, use_xyz=True, ret_grouped_xyz=False):
        super().__init__()
        self.use_xyz = use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            new_features = grouped_xyz
        if self.ret_grouped_xyz:
            return new_features, grouped_xyz
        return new_features
