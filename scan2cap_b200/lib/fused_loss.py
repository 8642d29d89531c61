"""VoteNet detection loss (vote + objectness + box + semantic terms of lib/loss_helper.py:24-187, 381-491) as ONE libs2c
launch that produces the loss terms, the label tensors AND the gradients (csrc/loss.cu), wrapped as an autograd
Function.  The framework formulation of the same arithmetic (lib/loss_helper.py of this package, used for CPU tensors
and as the comparison path of the tests) is ~200 kernels forward and ~300 backward per training step."""
import numpy as np
import torch
from torch.autograd import Function

from .._lib import call
from .pointnet2._ext import _guard, _stream

STAT_KEYS = ("det_loss", "vote_loss", "objectness_loss", "center_loss", "heading_cls_loss", "heading_reg_loss",
             "size_cls_loss", "size_reg_loss", "sem_cls_loss", "box_loss", "obj_acc", "pos_ratio", "neg_ratio")
_MEAN = {}


def _mean_size(mean_size_arr, device):
    key = (np.asarray(mean_size_arr, np.float32).tobytes(), str(device))
    if key not in _MEAN:
        _MEAN[key] = torch.from_numpy(np.ascontiguousarray(mean_size_arr, dtype=np.float32)).to(device)
    return _MEAN[key]


def available(data_dict):
    t = data_dict.get("_head_outputs")
    return isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32


class _DetectionLoss(Function):
    @staticmethod
    def forward(ctx, vote_xyz, net, center, agg_xyz, seed_xyz, seed_inds, vote_label, vote_label_mask, center_label,
                heading_class_label, heading_residual_label, size_class_label, size_residual_label, sem_cls_label,
                box_label_mask, mean_size, NH, NS, NC):
        B, S, _ = vote_xyz.shape
        K, W = net.shape[1], net.shape[2]
        G, N = center_label.shape[1], vote_label.shape[1]
        assert W == 5 + 2 * NH + 4 * NS + NC, (W, NH, NS, NC)
        dev = net.device
        c = lambda t: t if t.is_contiguous() else t.contiguous()
        vote_xyz, net, center, agg_xyz, seed_xyz = c(vote_xyz), c(net), c(center), c(agg_xyz), c(seed_xyz)
        if seed_inds.dtype != torch.int32 or seed_inds.stride(1) != 1:
            seed_inds = seed_inds.to(torch.int32).contiguous()
        labels = [c(t) for t in (vote_label, vote_label_mask, center_label[:, :, 0:3], heading_class_label,
                                 heading_residual_label, size_class_label, size_residual_label, sem_cls_label,
                                 box_label_mask)]
        assert labels[1].dtype == torch.int64 and labels[3].dtype == torch.int64 and labels[5].dtype == torch.int64
        stats = torch.empty(16, dtype=torch.float32, device=dev)
        obj_label = torch.empty((B, K), dtype=torch.int64, device=dev)
        obj_mask = torch.empty((B, K), dtype=torch.float32, device=dev)
        assign = torch.empty((B, K), dtype=torch.int64, device=dev)
        d_vote = torch.empty_like(vote_xyz)
        d_net = torch.empty_like(net)
        d_center = torch.empty_like(center)
        scratch = torch.empty(B * K + B * G, dtype=torch.int32, device=dev)
        with _guard(net):
            call("s2c_detection_loss", B, S, N, K, G, int(NH), int(NS), int(NC), vote_xyz.data_ptr(), seed_xyz.data_ptr(),
                 seed_inds.data_ptr(), seed_inds.stride(0), labels[0].data_ptr(), labels[1].data_ptr(),
                 agg_xyz.data_ptr(), net.data_ptr(), center.data_ptr(), labels[2].data_ptr(), labels[3].data_ptr(),
                 labels[4].data_ptr(), labels[5].data_ptr(), labels[6].data_ptr(), labels[7].data_ptr(),
                 labels[8].data_ptr(), mean_size.data_ptr(), stats.data_ptr(), obj_label.data_ptr(),
                 obj_mask.data_ptr(), assign.data_ptr(), d_vote.data_ptr(), d_net.data_ptr(), d_center.data_ptr(),
                 scratch.data_ptr(), _stream(net))
        ctx.save_for_backward(d_vote, d_net, d_center)
        ctx.mark_non_differentiable(stats, obj_label, obj_mask, assign)
        return stats[0].clone(), stats, obj_label, obj_mask, assign

    @staticmethod
    def backward(ctx, g, *unused):
        d_vote, d_net, d_center = ctx.saved_tensors
        return (g * d_vote, g * d_net, g * d_center) + (None,) * 16


def detection_loss(data_dict, config):
    """-> (det_loss [differentiable scalar: vote + 0.5 objectness + box + 0.1 sem_cls], dict of the individual terms /
    ratios, objectness_label (B,K) int64, objectness_mask (B,K) f32, object_assignment (B,K) int64)."""
    net = data_dict["_head_outputs"]
    det, stats, obj_label, obj_mask, assign = _DetectionLoss.apply(
        data_dict["vote_xyz"], net, data_dict["center"], data_dict["aggregated_vote_xyz"], data_dict["seed_xyz"],
        data_dict["seed_inds"], data_dict["vote_label"], data_dict["vote_label_mask"], data_dict["center_label"],
        data_dict["heading_class_label"], data_dict["heading_residual_label"], data_dict["size_class_label"],
        data_dict["size_residual_label"], data_dict["sem_cls_label"], data_dict["box_label_mask"],
        _mean_size(config.mean_size_arr, net.device), config.num_heading_bin, config.num_size_cluster, config.num_class)
    terms = {k: stats[i] for i, k in enumerate(STAT_KEYS)}
    return det, terms, obj_label, obj_mask, assign
