"""Mirror of models/mask_votenet.py (MaskVoteNet :221-293, its ProposalModule :134-218): the detector pre-training
variant -- same backbone / voting module, ONE proposal per scene aggregated over a 5 m ball with nsample = 512
(:145-153: a much larger group than any CapNet stage; same fused query+group and tensor-core MLP kernels), a head
without objectness / heading (3 + 4*NS + num_class outputs).  Same attribute names -> same checkpoint keys."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..lib.pointnet2 import fused_mlp
from ..lib.pointnet2.pointnet2_modules import PointnetSAModuleVotes
from .backbone_module import Pointnet2Backbone
from .voting_module import VotingModule


class ProposalModule(nn.Module):
    def __init__(self, num_class, num_heading_bin, num_size_cluster, mean_size_arr, num_proposal, sampling,
                 seed_feat_dim=256):
        super().__init__()
        self.num_class = num_class
        self.num_heading_bin = num_heading_bin
        self.num_size_cluster = num_size_cluster
        self.mean_size_arr = mean_size_arr
        self.num_proposal = num_proposal
        self.sampling = sampling
        self.seed_feat_dim = seed_feat_dim
        self.vote_aggregation = PointnetSAModuleVotes(npoint=self.num_proposal, radius=5, nsample=512,
                                                      mlp=[self.seed_feat_dim, 128, 128, 128], use_xyz=True,
                                                      normalize_xyz=True)
        self.proposal = nn.Sequential(
            nn.Conv1d(128, 128, 1, bias=False), nn.BatchNorm1d(128), nn.ReLU(),
            nn.Conv1d(128, 128, 1, bias=False), nn.BatchNorm1d(128), nn.ReLU(),
            nn.Conv1d(128, 3 + num_size_cluster * 4 + self.num_class, 1))
        self.register_buffer("_mean_size_f32", torch.from_numpy(mean_size_arr.astype(np.float32)), persistent=False)

    def forward(self, xyz, features, data_dict):
        xyz, features, fps_inds = self.vote_aggregation(xyz, features)
        data_dict["aggregated_vote_xyz"] = xyz
        data_dict["aggregated_vote_features"] = features.permute(0, 2, 1).contiguous()
        data_dict["aggregated_vote_inds"] = fps_inds
        head = self.proposal
        B, K = xyz.shape[0], xyz.shape[1]
        rows = data_dict["aggregated_vote_features"].reshape(B * K, -1)
        rows = fused_mlp.fused_mlp_maxpool(rows, rows.shape[1], B * K, 1, [(head[0], head[1]), (head[3], head[4])],
                                           self.training, capture=False)
        net = fused_mlp.linear_rows(rows, head[6].weight.view(head[6].weight.shape[0], -1), head[6].bias).view(B, K, -1)
        return self.decode_scores(net, data_dict)

    def decode_scores(self, net_transposed, data_dict):
        """net_transposed (B, num_proposal, 3 + 4*NS + num_class) (:196-218)."""
        B, K = net_transposed.shape[0], net_transposed.shape[1]
        NS = self.num_size_cluster
        data_dict["center"] = data_dict["aggregated_vote_xyz"] + net_transposed[:, :, 0:3]
        data_dict["size_scores"] = net_transposed[:, :, 3:3 + NS]
        srn = net_transposed[:, :, 3 + NS:3 + NS * 4].view(B, K, NS, 3)
        data_dict["size_residuals_normalized"] = srn
        data_dict["size_residuals"] = srn * self._mean_size_f32.unsqueeze(0).unsqueeze(0)
        data_dict["sem_cls_scores"] = net_transposed[:, :, 3 + NS * 4:]
        return data_dict


class MaskVoteNet(nn.Module):
    def __init__(self, num_class, num_heading_bin, num_size_cluster, mean_size_arr, input_feature_dim=0,
                 num_proposal=1, vote_factor=1, sampling="vote_fps"):
        super().__init__()
        self.num_class = num_class
        self.num_heading_bin = num_heading_bin
        self.num_size_cluster = num_size_cluster
        self.mean_size_arr = mean_size_arr
        assert mean_size_arr.shape[0] == self.num_size_cluster
        self.input_feature_dim = input_feature_dim
        self.num_proposal = num_proposal
        self.vote_factor = vote_factor
        self.sampling = sampling
        self.backbone_net = Pointnet2Backbone(input_feature_dim=self.input_feature_dim)
        self.vgen = VotingModule(self.vote_factor, 256)
        self.proposal = ProposalModule(num_class, num_heading_bin, num_size_cluster, mean_size_arr, num_proposal,
                                       sampling)

    def forward(self, data_dict):
        data_dict = self.backbone_net(data_dict)
        xyz = data_dict["fp2_xyz"]
        features = data_dict["fp2_features"]
        data_dict["seed_inds"] = data_dict["fp2_inds"]
        data_dict["seed_xyz"] = xyz
        data_dict["seed_features"] = features
        xyz, features = self.vgen(xyz, features)
        features_norm = torch.norm(features, p=2, dim=1)
        features = features.div(features_norm.unsqueeze(1))
        data_dict["vote_xyz"] = xyz
        data_dict["vote_features"] = features
        return self.proposal(xyz, features, data_dict)
