"""Mirror of models/backbone_module.py (Pointnet2Backbone :11-128): four set-abstraction layers and two
feature-propagation layers, same attribute names (sa1..sa4, fp1, fp2 -> checkpoint keys) and the same
``data_dict`` keys.  The point cloud's feature columns are fed to SA1 as they lie in ``point_clouds``
(point-major), so the reference's transpose().contiguous() copy (:68-71) does not happen."""
import torch.nn as nn

from ..lib.pointnet2.pointnet2_modules import PointnetSAModuleVotes, PointnetFPModule


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

    def _break_up_pc(self, pc):
        xyz = pc[..., :3].contiguous()
        # (B,C,N) VIEW of the point-major columns; PointnetSAModuleVotes consumes it without a copy
        features = pc[..., 3:].transpose(1, 2) if pc.size(-1) > 3 else None
        return xyz, features

    def forward(self, data_dict):
        pointcloud = data_dict["point_clouds"]
        xyz, features = self._break_up_pc(pointcloud)

        xyz, features, fps_inds = self.sa1(xyz, features)
        data_dict["sa1_inds"] = fps_inds
        data_dict["sa1_xyz"] = xyz
        data_dict["sa1_features"] = features

        xyz, features, fps_inds = self.sa2(xyz, features)
        data_dict["sa2_inds"] = fps_inds
        data_dict["sa2_xyz"] = xyz
        data_dict["sa2_features"] = features

        xyz, features, fps_inds = self.sa3(xyz, features)
        data_dict["sa3_xyz"] = xyz
        data_dict["sa3_features"] = features

        xyz, features, fps_inds = self.sa4(xyz, features)
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
