"""Mirror of models/proposal_module.py (ProposalModule :21-144): vote aggregation (SA layer with FPS on the
VOTE coordinates, npoint=num_proposal, r=0.3, nsample=16), the three-layer Conv1d head and the score / box
decoding.  Same attribute names (vote_aggregation, proposal.{0..6}) and ``data_dict`` keys.

decode_pred_box: the reference copies five tensors to the host, builds the corners with numpy in float64
and copies them back (:80-103, one D2H + one H2D sync per step).  Here the same float64 arithmetic runs on
the device (utils/box_util.axis_aligned_corners), bit-identical and sync-free."""
import numpy as np
import torch
import torch.nn as nn

from ..data.scannet.model_util_scannet import ScannetDatasetConfig

from ..lib.pointnet2 import fused_mlp
from ..lib.pointnet2.pointnet2_modules import PointnetSAModuleVotes
from ..utils.box_util import axis_aligned_corners

DC = ScannetDatasetConfig()


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
        self.vote_aggregation = PointnetSAModuleVotes(npoint=self.num_proposal, radius=0.3, nsample=16,
                                                      mlp=[self.seed_feat_dim, 128, 128, 128], use_xyz=True,
                                                      normalize_xyz=True)
        self.proposal = nn.Sequential(
            nn.Conv1d(128, 128, 1, bias=False), nn.BatchNorm1d(128), nn.ReLU(),
            nn.Conv1d(128, 128, 1, bias=False), nn.BatchNorm1d(128), nn.ReLU(),
            nn.Conv1d(128, 2 + 3 + num_heading_bin * 2 + num_size_cluster * 4 + self.num_class, 1))
        # constants kept on the module's device (not persistent: the reference state dict has no such keys)
        self.register_buffer("_mean_size_f32", torch.from_numpy(mean_size_arr.astype(np.float32)), persistent=False)
        self.register_buffer("_mean_size_f64", torch.from_numpy(DC.mean_size_arr.astype(np.float64)), persistent=False)

    def forward(self, xyz, features, data_dict):
        xyz, features, fps_inds = self.vote_aggregation(xyz, features)
        data_dict["aggregated_vote_xyz"] = xyz
        data_dict["aggregated_vote_features"] = features.permute(0, 2, 1).contiguous()
        data_dict["aggregated_vote_inds"] = fps_inds
        # the Conv1d/BN/ReLU head on the point-major (B*K, 128) rows (same arithmetic, no transposes): the two BatchNorm
        # layers as a fused stack on the tensor-core layer kernels, the last Conv1d (bias, no BatchNorm) as linear_rows
        head = self.proposal
        B, K = xyz.shape[0], xyz.shape[1]
        rows = data_dict["aggregated_vote_features"].reshape(B * K, -1)
        rows = fused_mlp.fused_mlp_maxpool(rows, rows.shape[1], B * K, 1, [(head[0], head[1]), (head[3], head[4])],
                                           self.training, capture=False)
        net = fused_mlp.linear_rows(rows, head[6].weight.view(head[6].weight.shape[0], -1), head[6].bias)
        net = net.view(B, K, -1).transpose(2, 1)  # (B,97,K) view
        return self.decode_scores(net, data_dict, self.num_class, self.num_heading_bin, self.num_size_cluster,
                                  self.mean_size_arr)

    def decode_pred_box(self, data_dict):
        """(B,K,8,3) float64 corners; arithmetic of DC.param2obb_batch + get_3d_box_batch in float64."""
        center = data_dict["center"].detach().double()
        size_class = torch.argmax(data_dict["size_scores"], -1)  # (B,K)
        size_residual = torch.gather(data_dict["size_residuals"].detach(), 2,
                                     size_class.view(*size_class.shape, 1, 1).expand(-1, -1, 1, 3)).squeeze(2)
        box_size = self._mean_size_f64[size_class] + size_residual.double()
        return axis_aligned_corners(box_size, center)

    def decode_scores(self, net, data_dict, num_class, num_heading_bin, num_size_cluster, mean_size_arr):
        net_transposed = net.transpose(2, 1).contiguous()
        batch_size, num_proposal = net_transposed.shape[0], net_transposed.shape[1]
        NH, NS = num_heading_bin, num_size_cluster
        objectness_scores = net_transposed[:, :, 0:2]
        center = data_dict["aggregated_vote_xyz"] + net_transposed[:, :, 2:5]
        heading_scores = net_transposed[:, :, 5:5 + NH]
        heading_residuals_normalized = net_transposed[:, :, 5 + NH:5 + NH * 2]
        size_scores = net_transposed[:, :, 5 + NH * 2:5 + NH * 2 + NS]
        size_residuals_normalized = net_transposed[:, :, 5 + NH * 2 + NS:5 + NH * 2 + NS * 4].view(
            [batch_size, num_proposal, NS, 3])
        sem_cls_scores = net_transposed[:, :, 5 + NH * 2 + NS * 4:]

        data_dict["_head_outputs"] = net_transposed   # (B,K,97) packed head outputs: input of the fused loss kernel
        data_dict["objectness_scores"] = objectness_scores
        data_dict["center"] = center
        data_dict["heading_scores"] = heading_scores
        data_dict["heading_residuals_normalized"] = heading_residuals_normalized
        data_dict["heading_residuals"] = heading_residuals_normalized * (np.pi / NH)
        data_dict["size_scores"] = size_scores
        data_dict["size_residuals_normalized"] = size_residuals_normalized
        data_dict["size_residuals"] = size_residuals_normalized * self._mean_size_f32.unsqueeze(0).unsqueeze(0)
        data_dict["sem_cls_scores"] = sem_cls_scores

        data_dict["bbox_corner"] = self.decode_pred_box(data_dict)
        data_dict["bbox_feature"] = data_dict["aggregated_vote_features"]
        data_dict["bbox_mask"] = objectness_scores.argmax(-1)
        data_dict["bbox_sems"] = sem_cls_scores.argmax(-1)
        data_dict["sem_cls"] = sem_cls_scores.argmax(-1)
        return data_dict
