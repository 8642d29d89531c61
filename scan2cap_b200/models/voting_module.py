"""Mirror of models/voting_module.py (VotingModule :9-60): Conv1d(256)->BN->ReLU x2 -> Conv1d(3+256) gives a
per-seed xyz offset and a residual feature.  Same attribute names (conv1..3, bn1..2 -> checkpoint keys).
The pointwise convolutions run on the point-major (B*S, C) row matrix (no transposes / copies)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..lib.pointnet2.pointnet2_modules import bn_rows, conv1x1_rows, point_major


class VotingModule(nn.Module):
    def __init__(self, vote_factor, seed_feature_dim):
        super().__init__()
        self.vote_factor = vote_factor
        self.in_dim = seed_feature_dim
        self.out_dim = self.in_dim  # residual feature: in_dim == out_dim
        self.conv1 = torch.nn.Conv1d(self.in_dim, self.in_dim, 1)
        self.conv2 = torch.nn.Conv1d(self.in_dim, self.in_dim, 1)
        self.conv3 = torch.nn.Conv1d(self.in_dim, (3 + self.out_dim) * self.vote_factor, 1)
        self.bn1 = torch.nn.BatchNorm1d(self.in_dim)
        self.bn2 = torch.nn.BatchNorm1d(self.in_dim)

    def forward(self, seed_xyz, seed_features):
        """seed_xyz (B,S,3), seed_features (B,C,S) -> vote_xyz (B,S*vf,3), vote_features (B,C,S*vf)."""
        batch_size, num_seed = seed_xyz.shape[0], seed_xyz.shape[1]
        num_vote = num_seed * self.vote_factor
        seed_pm = point_major(seed_features)  # (B,S,C)
        rows = seed_pm.reshape(batch_size * num_seed, self.in_dim)
        net = F.relu(bn_rows(conv1x1_rows(rows, self.conv1), self.bn1, self.training))
        net = F.relu(bn_rows(conv1x1_rows(net, self.conv2), self.bn2, self.training))
        net = conv1x1_rows(net, self.conv3).view(batch_size, num_seed, self.vote_factor, 3 + self.out_dim)
        vote_xyz = (seed_xyz.unsqueeze(2) + net[:, :, :, 0:3]).contiguous().view(batch_size, num_vote, 3)
        vote_features = (seed_pm.unsqueeze(2) + net[:, :, :, 3:]).contiguous().view(batch_size, num_vote, self.out_dim)
        return vote_xyz, vote_features.transpose(2, 1)  # (B,C,S*vf) view over point-major storage
