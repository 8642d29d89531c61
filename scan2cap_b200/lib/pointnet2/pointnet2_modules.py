"""Mirror of the reference's lib/pointnet2/pointnet2_modules.py: PointnetSAModuleVotes (:164-272) and
PointnetFPModule (:356-416) with identical constructor signatures, forward signatures / return values and
state-dict keys (child ``mlp_module`` / ``mlp`` holding ``layer{i}.conv`` and ``layer{i}.bn.bn``).

What differs is how the work is issued on the B200:
  * FPS also emits the sampled coordinates (no transpose + gather_points round trip);
  * ball query + both gathers + centre subtraction + 1/radius + concat are ONE kernel that writes the grouped
    tensor channels-last, i.e. as the row-major (B*npoint*nsample, 3+C) matrix the per-group MLP multiplies;
  * per-point features travel between layers point-major ((B,N,C) storage behind the (B,C,N) view the
    reference API promises), so no layer ever transposes or copies a feature map.
The MSG / LFP variants of the reference (:78, :127, :274, :418) are never instantiated by CapNet and are not
provided.
"""
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

import os

from . import _ext
from . import fused_mlp
from . import pointnet2_utils
from . import pytorch_utils as pt_utils

# S2C_FUSED_MLP=0 selects library GEMM + BatchNorm + ReLU kernels instead of the tcgen05 path (A/B comparison)
USE_FUSED_MLP = os.environ.get("S2C_FUSED_MLP", "1") != "0"


def point_major(features):
    """(B,C,N) tensor -> (B,N,C) view with unit channel stride (copying only if the storage is channel-major)."""
    f = features.transpose(1, 2)
    return f if f.stride(2) == 1 and f.stride(0) == f.shape[1] * f.stride(1) else f.contiguous()


def bn_rows(x, bn, training):
    """BatchNorm{1d,2d} of a row-major (R, C) matrix: statistics over the R rows (every (scene, point[, sample]))."""
    if training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    mom = bn.momentum if bn.momentum is not None else 0.0
    return F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias,
                        training or not bn.track_running_stats, mom, bn.eps)


def conv1x1_rows(x, conv):
    """Pointwise Conv1d/Conv2d applied to a row-major (R, Cin) matrix."""
    return F.linear(x, conv.weight.view(conv.weight.shape[0], -1), conv.bias)


def shared_mlp_rows(rows, layers, training):
    """conv1x1(no bias) -> BatchNorm -> ReLU stack on a row-major (R, Cin) matrix (R = every (scene, group,
    sample) triple), the arithmetic of SharedMLP on a (B,C,npoint,nsample) tensor (pytorch_utils.py:11-36,
    88-120).  BatchNorm statistics over R rows == BatchNorm2d statistics over (B, npoint, nsample)."""
    x = rows
    for conv, bn in layers:
        x = conv1x1_rows(x, conv)
        if bn is not None:
            x = bn_rows(x, bn, training)
        x = F.relu_(x)
    return x


class PointnetSAModuleVotes(nn.Module):
    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None,
                 bn: bool = True, use_xyz: bool = True, pooling: str = "max", sigma: float = None,
                 normalize_xyz: bool = False, sample_uniformly: bool = False, ret_unique_cnt: bool = False):
        super().__init__()
        self.npoint = npoint
        self.radius = radius
        self.nsample = nsample
        self.pooling = pooling
        self.mlp_module = None
        self.use_xyz = use_xyz
        self.sigma = sigma
        if self.sigma is None and self.radius is not None:
            self.sigma = self.radius / 2
        self.normalize_xyz = normalize_xyz
        self.ret_unique_cnt = ret_unique_cnt
        if npoint is not None:
            self.grouper = pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz, ret_grouped_xyz=True,
                                                         normalize_xyz=normalize_xyz,
                                                         sample_uniformly=sample_uniformly,
                                                         ret_unique_cnt=ret_unique_cnt)
        else:
            self.grouper = pointnet2_utils.GroupAll(use_xyz, ret_grouped_xyz=True)
        mlp_spec = mlp
        if use_xyz and len(mlp_spec) > 0:
            mlp_spec[0] += 3  # (the reference mutates the caller's list too: pointnet2_modules.py:204-206)
        self.mlp_module = pt_utils.SharedMLP(mlp_spec, bn=bn)

    def forward(self, xyz: torch.Tensor, features: torch.Tensor = None, inds: torch.Tensor = None,
                sampled_xyz: torch.Tensor = None):
        """xyz (B,N,3), features (B,C,N), inds (B,npoint) -> new_xyz (B,npoint,3), new_features (B,C',npoint),
        inds (B,npoint) int32.  sampled_xyz (extension): xyz[inds] when the caller already has it (no gradient)."""
        layers = self.mlp_module.layer_params()
        fast = (self.npoint is not None and self.use_xyz and self.pooling == "max" and layers is not None
                and not self.ret_unique_cnt)
        if not fast:
            return self._forward_generic(xyz, features, inds)
        xyz = xyz.contiguous()
        if inds is None:
            inds, new_xyz = _ext.furthest_point_sampling_with_xyz(xyz.detach(), self.npoint)
            if xyz.requires_grad:  # vote aggregation: the sampled coordinates carry gradient
                new_xyz = torch.gather(xyz, 1, inds.long().unsqueeze(-1).expand(-1, -1, 3))
        elif sampled_xyz is not None and not xyz.requires_grad:
            assert inds.shape[1] == self.npoint
            new_xyz = sampled_xyz
        else:
            assert inds.shape[1] == self.npoint
            new_xyz = torch.gather(xyz, 1, inds.long().unsqueeze(-1).expand(-1, -1, 3))
        feats_pm = point_major(features) if features is not None else None
        use_fused = USE_FUSED_MLP and fused_mlp.fusable(layers)
        grouped, _ = pointnet2_utils.query_and_group(xyz, new_xyz, feats_pm, self.radius, self.nsample,
                                                     self.normalize_xyz, True, True, use_fused)
        B, C, M, ns = grouped.shape
        if use_fused:
            # rows of the 16-byte aligned channels-last buffer the kernel wrote: (R, Cp) = [xyz, 0 | features | pad]
            Cin = 3 + (feats_pm.shape[2] if feats_pm is not None else 0)
            rows = grouped.permute(0, 2, 3, 1).reshape(B * M * ns, C)
            pooled = fused_mlp.fused_mlp_maxpool(rows, Cin, B * M, ns, layers, self.training, xyz_gap=True,
                                                 need_xyz_grad=xyz.requires_grad or new_xyz.requires_grad
                                                 ).view(B, M, -1)
        else:
            rows = grouped.permute(0, 2, 3, 1).reshape(B * M * ns, C)  # a view: the kernel wrote channels-last
            out = shared_mlp_rows(rows, layers, self.training)
            pooled = out.view(B, M, ns, -1).amax(dim=2)  # (B, M, C') point-major
        return new_xyz, pooled.transpose(1, 2), inds

    def _forward_generic(self, xyz, features, inds):
        """Literal path for the options CapNet never uses (avg / rbf pooling, GroupAll, use_xyz=False)."""
        xyz_flipped = xyz.transpose(1, 2).contiguous()
        if inds is None and self.npoint is not None:
            inds = pointnet2_utils.furthest_point_sample(xyz, self.npoint)
        new_xyz = (pointnet2_utils.gather_operation(xyz_flipped, inds).transpose(1, 2).contiguous()
                   if self.npoint is not None else None)
        grouped_features, grouped_xyz = self.grouper(xyz, new_xyz, features.contiguous() if features is not None else None)
        new_features = self.mlp_module(grouped_features)
        if self.pooling == "max":
            new_features = F.max_pool2d(new_features, kernel_size=[1, new_features.size(3)])
        elif self.pooling == "avg":
            new_features = F.avg_pool2d(new_features, kernel_size=[1, new_features.size(3)])
        elif self.pooling == "rbf":
            rbf = torch.exp(-1 * grouped_xyz.pow(2).sum(1, keepdim=False) / (self.sigma ** 2) / 2)
            new_features = torch.sum(new_features * rbf.unsqueeze(1), -1, keepdim=True) / float(self.nsample)
        return new_xyz, new_features.squeeze(-1), inds


class PointnetFPModule(nn.Module):
    def __init__(self, *, mlp: List[int], bn: bool = True):
        super().__init__()
        self.mlp = pt_utils.SharedMLP(mlp, bn=bn)

    def forward(self, unknown, known, unknow_feats, known_feats):
        """unknown (B,n,3), known (B,m,3), unknow_feats (B,C1,n), known_feats (B,C2,m) -> (B,mlp[-1],n)."""
        if known is not None:
            dist, idx = pointnet2_utils.three_nn(unknown.contiguous(), known.contiguous())
            dist_recip = 1.0 / (dist + 1e-8)
            norm = torch.sum(dist_recip, dim=2, keepdim=True)
            weight = dist_recip / norm
            interpolated_feats = pointnet2_utils.three_interpolate(known_feats.contiguous(), idx, weight)
        else:
            interpolated_feats = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))
        layers = self.mlp.layer_params()
        if layers is None:
            new_features = (torch.cat([interpolated_feats, unknow_feats], dim=1) if unknow_feats is not None
                            else interpolated_feats)
            return self.mlp(new_features.unsqueeze(-1)).squeeze(-1)
        parts = [interpolated_feats.transpose(1, 2)]
        if unknow_feats is not None:
            parts.append(unknow_feats.transpose(1, 2))
        rows = torch.cat(parts, dim=2)  # (B, n, C2+C1) point-major
        B, n, C = rows.shape
        if USE_FUSED_MLP and fused_mlp.fusable(layers):
            out = fused_mlp.fused_mlp_maxpool(rows.reshape(B * n, C), C, B * n, 1, layers, self.training)
        else:
            out = shared_mlp_rows(rows.reshape(B * n, C), layers, self.training)
        return out.view(B, n, -1).transpose(1, 2)
