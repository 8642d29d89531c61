"""Mirror of models/backbone_module.py (Pointnet2Backbone :11-128): four set-abstraction layers and two
feature-propagation layers, same attribute names (sa1..sa4, fp1, fp2 -> checkpoint keys) and the same
``data_dict`` keys.  The point cloud's feature columns are fed to SA1 as they lie in ``point_clouds``
(point-major), so the reference's transpose().contiguous() copy (:68-71) does not happen."""
import os

import torch
import torch.nn as nn

from ..lib.pointnet2 import _ext
from ..lib.pointnet2.pointnet2_modules import PointnetSAModuleVotes, PointnetFPModule

# The sampling of SA2..SA4 depends on coordinates only (each level samples the previous level's samples), so it can
# run on a second stream next to SA1's grouping + MLP instead of between the levels (S2C_SAMPLE_AHEAD=0: in sequence)
SAMPLE_AHEAD = True
_SIDE_STREAMS = {}


def padded_point_clouds(pc):
    """A copy of point_clouds (B,N,3+C) whose rows are [pad, x, y, z | C features | pad to 4 floats]: the returned
    (B,N,3+C) VIEW has the same values, 16-byte aligned feature rows and a row stride that is a multiple of 16 bytes --
    what the TMA gather of the fused query+group kernel (and its float4 path for narrow rows) needs."""
    B, N, F = pc.shape
    stride = 4 + (F - 3 + 3) // 4 * 4
    buf = torch.empty((B, N, stride), dtype=pc.dtype, device=pc.device)
    view = buf[..., 1:1 + F]
    view.copy_(pc)
    return view


def padded_point_clouds_like(shape, dtype, device):
    """An uninitialised tensor of `shape` = (B,N,3+C) in the layout of padded_point_clouds()."""
    B, N, F = shape
    stride = 4 + (F - 3 + 3) // 4 * 4
    return torch.empty((B, N, stride), dtype=dtype, device=device)[..., 1:1 + F]


class Pointnet2Backbone(nn.Module):
    def __init__(self, input_feature_dim=0):
        super().__init__()
        self.input_feature_dim = input_feature_dim
        self.sa1 = PointnetSAModuleVotes(npoint=2048, radius=0.2, nsample=64,
                                         mlp=[input_feature_dim, 64, 64, 128], use_xyz=True, normalize_xyz=True)
        self.sa2 = PointnetSAModuleVotes(npoint=1024, radius=0.4, nsample=32,
                                         mlp=[128, 128, 128, 256], use_xyz=True, normalize_xyz=True)
        self.sa3 = PointnetSAModuleVotes(npoint=512, radius=0.8, nsample=16,
                                         mlp=[256, 128, 128, 256], use_xyz=True, normalize_xyz=True)
        self.sa4 = PointnetSAModuleVotes(npoint=256, radius=1.2, nsample=16,
                                         mlp=[256, 128, 128, 256], use_xyz=True, normalize_xyz=True)
        self.fp1 = PointnetFPModule(mlp=[256 + 256, 256, 256])
        self.fp2 = PointnetFPModule(mlp=[256 + 256, 256, 256])

    @staticmethod
    def _side_stream(device):
        key = (device.type, device.index)
        if key not in _SIDE_STREAMS:
            _SIDE_STREAMS[key] = torch.cuda.Stream(device)
        return _SIDE_STREAMS[key]

    def _break_up_pc(self, pc):
        xyz = pc[..., :3].contiguous()
        # (B,C,N) VIEW of the point-major columns; PointnetSAModuleVotes consumes it without a copy
        if pc.size(-1) <= 3:
            return xyz, None
        feats = pc[..., 3:]
        C = feats.shape[-1]
        if C >= 32 and pc.is_cuda and (feats.data_ptr() % 16 != 0 or feats.stride(1) % 4 != 0):
            # wide feature rows that are not 16-byte aligned (a contiguous (B,N,3+C) point cloud never is): repack them once
            # into the aligned layout of padded_point_clouds() so SA1's gather runs on the TMA unit (65 % of the HBM
            # roofline instead of 33 % from 4-byte loads).  engine.TrainStep keeps its static input buffer in that layout,
            # where this copy does not happen.
            feats = padded_point_clouds(pc)[..., 3:]
        return xyz, feats.transpose(1, 2)

    def sample_indices(self, xyz):
        """FPS of all four levels (40000 -> 2048 -> 1024 -> 512 -> 256 for the default sizes) from coordinates alone:
        [(inds (B,m) int32, xyz (B,m,3))] x 4, enqueued on the current stream.  The indices depend on nothing but
        point_clouds[..., :3], so a training loop can compute them for batch i+1 while batch i is in flight
        (engine.TrainStep.prefetch) and pass them in as data_dict["fps_precomputed"]."""
        out = []
        cur = xyz.contiguous()
        for m in (self.sa1.npoint, self.sa2.npoint, self.sa3.npoint, self.sa4.npoint):
            inds, cur = _ext.furthest_point_sampling_with_xyz(cur, m)
            out.append((inds, cur))
        return out

    def sa1_grid(self, xyz, out=None):
        """The uniform ball-query grid SA1 searches (depends only on the coordinates and SA1's radius): like the sampling
        indices it can be built for batch i+1 while batch i is in flight and passed in as data_dict["sa1_grid"]."""
        return _ext.ball_query_grid_build(xyz.contiguous(), self.sa1.radius, out=out)

    def forward(self, data_dict):
        pointcloud = data_dict["point_clouds"]
        xyz, features = self._break_up_pc(pointcloud)

        ahead = [(None, None)] * 3
        pre = data_dict.get("fps_precomputed")
        if pre is not None:
            # sampled ahead of time (same kernels, same result): the 2 ms serial FPS chain is off the step's critical path
            (inds1, xyz1), ahead = pre[0], list(pre[1:])
            xyz, features, fps_inds = self.sa1(xyz, features, inds1, sampled_xyz=xyz1, grid=data_dict.get("sa1_grid"))
        elif SAMPLE_AHEAD and xyz.is_cuda and not xyz.requires_grad:
            # FPS of all four levels up front: level 1 on this stream, levels 2-4 (8 CTAs each) on a side stream that
            # overlaps SA1's grouping and MLP; the streams join before SA2 (a fork/join the CUDA-graph capture keeps)
            inds1, xyz1 = _ext.furthest_point_sampling_with_xyz(xyz, self.sa1.npoint)
            main = torch.cuda.current_stream(xyz.device)
            side = self._side_stream(xyz.device)
            side.wait_stream(main)
            ahead = _ext.furthest_point_sampling_chain(xyz1, [self.sa2.npoint, self.sa3.npoint, self.sa4.npoint], side)
            xyz, features, fps_inds = self.sa1(xyz, features, inds1, sampled_xyz=xyz1)
            main.wait_stream(side)
        else:
            xyz, features, fps_inds = self.sa1(xyz, features)
        data_dict["sa1_inds"] = fps_inds
        data_dict["sa1_xyz"] = xyz
        data_dict["sa1_features"] = features

        xyz, features, fps_inds = self.sa2(xyz, features, ahead[0][0], sampled_xyz=ahead[0][1])
        data_dict["sa2_inds"] = fps_inds
        data_dict["sa2_xyz"] = xyz
        data_dict["sa2_features"] = features

        xyz, features, fps_inds = self.sa3(xyz, features, ahead[1][0], sampled_xyz=ahead[1][1])
        data_dict["sa3_xyz"] = xyz
        data_dict["sa3_features"] = features

        xyz, features, fps_inds = self.sa4(xyz, features, ahead[2][0], sampled_xyz=ahead[2][1])
        data_dict["sa4_xyz"] = xyz
        data_dict["sa4_features"] = features

        features = self.fp1(data_dict["sa3_xyz"], data_dict["sa4_xyz"], data_dict["sa3_features"],
                            data_dict["sa4_features"])
        features = self.fp2(data_dict["sa2_xyz"], data_dict["sa3_xyz"], data_dict["sa2_features"], features)
        data_dict["fp2_features"] = features
        data_dict["fp2_xyz"] = data_dict["sa2_xyz"]
        num_seed = data_dict["fp2_xyz"].shape[1]
        data_dict["fp2_inds"] = data_dict["sa1_inds"][:, 0:num_seed]
        return data_dict
