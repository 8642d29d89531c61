"""Mirror of models/voting_module.py (VotingModule :9-60): Conv1d(256)->BN->ReLU x2 -> Conv1d(3+256) gives a
per-seed xyz offset and a residual feature.  Same attribute names (conv1..3, bn1..2 -> checkpoint keys).
The pointwise convolutions run on the point-major (B*S, C) row matrix (no transposes / copies), on the tensor-core layer
kernels of libs2c: conv1/bn1/relu/conv2/bn2/relu as a fused two-layer stack (train-mode BatchNorm from the GEMM epilogue's
statistics; the conv biases in front of BatchNorm are folded, see fused_mlp._fold_conv_bias), conv3 as linear_rows."""
import torch
import torch.nn as nn

from ..lib.pointnet2 import fused_mlp
from ..lib.pointnet2.pointnet2_modules import point_major


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
        if not rows.is_contiguous():
            rows = rows.contiguous()
        net = fused_mlp.fused_mlp_maxpool(rows, self.in_dim, rows.shape[0], 1,
                                          [(self.conv1, self.bn1), (self.conv2, self.bn2)], self.training, capture=False)
        net = fused_mlp.linear_rows(net, self.conv3.weight.view(self.conv3.weight.shape[0], -1), self.conv3.bias)
        net = net.view(batch_size, num_seed, self.vote_factor, 3 + self.out_dim)
        vote_xyz = (seed_xyz.unsqueeze(2) + net[:, :, :, 0:3]).contiguous().view(batch_size, num_vote, 3)
        vote_features = (seed_pm.unsqueeze(2) + net[:, :, :, 3:]).contiguous().view(batch_size, num_vote, self.out_dim)
        return vote_xyz, vote_features.transpose(2, 1)  # (B,C,S*vf) view over point-major storage
