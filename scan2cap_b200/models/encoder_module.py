"""Mirror of models/encoder_module.py (PointnetEncoder :11-202): four set-abstraction layers, a global max over the 256
remaining points, Linear(256,128)+ReLU and an 18-way classifier; whole_scene=True encodes every masked object of a
scene ((B, num_bboxes, N, 3+C) input).  The reference loops over scenes and over chunks of `batch_size` objects
(:86-139); here all valid objects of a scene run as one batch (same arithmetic per object in eval mode; in training
mode BatchNorm statistics are taken over the scene's objects instead of per chunk)."""
import torch
import torch.nn as nn

from ..lib.pointnet2.pointnet2_modules import PointnetSAModuleVotes


class PointnetEncoder(nn.Module):
    def __init__(self, input_feature_dim=0, num_classes=18, whole_scene=False):
        super().__init__()
        self.input_feature_dim = input_feature_dim
        self.num_classes = num_classes
        self.whole_scene = whole_scene
        self.sa1 = PointnetSAModuleVotes(npoint=2048, radius=0.2, nsample=64, mlp=[input_feature_dim, 64, 64, 128],
                                         use_xyz=True, normalize_xyz=True)
        self.sa2 = PointnetSAModuleVotes(npoint=1024, radius=0.4, nsample=32, mlp=[128, 128, 128, 256],
                                         use_xyz=True, normalize_xyz=True)
        self.sa3 = PointnetSAModuleVotes(npoint=512, radius=0.8, nsample=16, mlp=[256, 128, 128, 256],
                                         use_xyz=True, normalize_xyz=True)
        self.sa4 = PointnetSAModuleVotes(npoint=256, radius=1.2, nsample=16, mlp=[256, 128, 128, 256],
                                         use_xyz=True, normalize_xyz=True)
        self.map = nn.Sequential(nn.Linear(256, 128), nn.ReLU())
        self.classifier = nn.Linear(128, num_classes)

    def _break_up_pc(self, pc):
        xyz = pc[..., :3].contiguous()
        features = pc[..., 3:].transpose(1, 2) if pc.size(-1) > 3 else None
        return xyz, features

    def _encode(self, pc, data_dict):
        xyz, features = self._break_up_pc(pc)
        for i, sa in enumerate((self.sa1, self.sa2, self.sa3, self.sa4), 1):
            xyz, features, fps_inds = sa(xyz, features)
            data_dict["sa%d_inds" % i] = fps_inds
            data_dict["sa%d_xyz" % i] = xyz
            data_dict["sa%d_features" % i] = features
        features = self.map(features.max(-1)[0])
        return features, self.classifier(features)

    def forward(self, data_dict):
        pointcloud = data_dict["point_clouds"]
        if not self.whole_scene:
            data_dict["enc_features"], data_dict["enc_preds"] = self._encode(pointcloud, data_dict)
            return data_dict
        object_masks = data_dict["target_masks"]  # (B, num_bboxes)
        B, num_bboxes = pointcloud.shape[0], pointcloud.shape[1]
        enc_features = pointcloud.new_zeros(B, num_bboxes, 128)
        enc_preds = pointcloud.new_zeros(B, num_bboxes, self.num_classes)
        for i in range(B):
            keep = (object_masks[i] == 1).nonzero(as_tuple=True)[0]
            if keep.numel() == 0:
                continue
            feats, preds = self._encode(pointcloud[i].index_select(0, keep), data_dict)
            enc_features[i] = enc_features[i].index_copy(0, keep, feats)
            enc_preds[i] = enc_preds[i].index_copy(0, keep, preds)
        data_dict["enc_features"] = enc_features
        data_dict["enc_preds"] = enc_preds
        return data_dict
