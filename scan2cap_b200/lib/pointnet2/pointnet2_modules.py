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

from . import _ext
from . import fused_mlp
from . import pointnet2_utils
from . import pytorch_utils as pt_utils

def point_major(features):
    """(B,C,N) tensor -> (B,N,C) view with unit channel stride (copying only if the storage is channel-major)."""
    f = features.transpose(1, 2)
    return f if f.stride(2) == 1 and f.stride(0) == f.shape[1] * f.stride(1) else f.contiguous()


def _require_fusable(layers, what):
    """The grouped MLP runs on the tcgen05 kernels only; there is no library-kernel alternative in the product."""
    if not fused_mlp.fusable(layers):
        raise RuntimeError("%s: the shared MLP must be [1x1 conv without bias -> affine BatchNorm -> ReLU] layers with "
                           "output widths that are multiples of 16 and <= 256 (tensor-core kernels of libs2c)" % what)


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
                sampled_xyz: torch.Tensor = None, grid: torch.Tensor = None):
        """xyz (B,N,3), features (B,C,N), inds (B,npoint) -> new_xyz (B,npoint,3), new_features (B,C',npoint),
        inds (B,npoint) int32.  Extensions: sampled_xyz = xyz[inds] when the caller already has it (no gradient);
        grid = the ball-query grid of (xyz, self.radius) built ahead of time (_ext.ball_query_grid_build)."""
        layers = self.mlp_module.layer_params()
        if self.npoint is None or not self.use_xyz or self.ret_unique_cnt or self.pooling not in ("max", "avg", "rbf"):
            # GroupAll raises in the reference itself (pointnet2_utils.py:387-390 vs :422);
            # use_xyz=False / ret_unique_cnt are never used by the CapNet / MaskVoteNet / encoder stacks
            raise NotImplementedError("PointnetSAModuleVotes: npoint=None, use_xyz=False and ret_unique_cnt are not "
                                      "provided by the fused query+group kernel")
        _require_fusable(layers, "PointnetSAModuleVotes")
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
        grouped, _ = pointnet2_utils.query_and_group(xyz, new_xyz, feats_pm, self.radius, self.nsample,
                                                     self.normalize_xyz, True, True, True, grid=grid)
        B, C, M, ns = grouped.shape
        # rows of the 16-byte aligned channels-last buffer the kernel wrote: (R, Cp) = [xyz, 0 | features | pad]
        Cin = 3 + (feats_pm.shape[2] if feats_pm is not None else 0)
        rows = grouped.permute(0, 2, 3, 1).reshape(B * M * ns, C)
        need_xyz_grad = xyz.requires_grad or new_xyz.requires_grad
        if self.pooling == "max":
            pooled = fused_mlp.fused_mlp_maxpool(rows, Cin, B * M, ns, layers, self.training, xyz_gap=True,
                                                 need_xyz_grad=need_xyz_grad).view(B, M, -1)
        else:
            # avg / rbf pooling (pointnet2_modules.py:258-266; never used by CapNet): same kernels without the pooling
            # stage (groups of one row), then the weighted mean over the nsample rows of each group
            act = fused_mlp.fused_mlp_maxpool(rows, Cin, B * M * ns, 1, layers, self.training, xyz_gap=True,
                                              need_xyz_grad=need_xyz_grad).view(B, M, ns, -1)
            if self.pooling == "avg":
                pooled = act.mean(dim=2)
            else:
                gxyz = rows[:, :3].view(B, M, ns, 3)  # grouped_xyz as QueryAndGroup returns it (centred [, / radius])
                rbf = torch.exp(-1 * gxyz.pow(2).sum(-1) / (self.sigma ** 2) / 2)
                pooled = torch.sum(act * rbf.unsqueeze(-1), 2) / float(self.nsample)
        return new_xyz, pooled.transpose(1, 2), inds


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
        _require_fusable(layers, "PointnetFPModule")
        out = fused_mlp.fused_mlp_maxpool(rows.reshape(B * n, C), C, B * n, 1, layers, self.training)
        return out.view(B, n, -1).transpose(1, 2)
