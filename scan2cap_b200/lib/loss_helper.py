"""Mirror of lib/loss_helper.py (get_scene_cap_loss :381-491 and the losses it calls): VoteNet vote /
objectness / box / semantic losses, caption cross-entropy, relative-orientation and distance losses of the
graph edges.  Plain PyTorch, as in the reference -- it is here to drive the backward pass of the hot path
 -- but written without host synchronisation: boolean-mask indexing, per-scene
Python loops (:272-299, :336-350) and Python-side branches on device values are replaced by masked
reductions over fixed shapes, so a whole training step can be captured in a CUDA graph.  Same keys written to
``data_dict``, same weighting (:472-487)."""
import numpy as np
import torch
import torch.nn.functional as F

from ..utils.nn_distance import nn_distance, huber_loss
from . import fused_loss
from .config import CONF

FAR_THRESHOLD = 0.6
NEAR_THRESHOLD = 0.3
GT_VOTE_FACTOR = 3
OBJECTNESS_CLS_WEIGHTS = [0.2, 0.8]

_CONST = {}


def _const(name, make, device):
    """Device-resident constants, created once per device (no H2D copy inside a captured step)."""
    key = (name, str(device))
    if key not in _CONST:
        _CONST[key] = make().to(device)
    return _CONST[key]


def compute_vote_loss(data_dict):
    batch_size, num_seed = data_dict["seed_xyz"].shape[0], data_dict["seed_xyz"].shape[1]
    vote_xyz = data_dict["vote_xyz"]
    seed_inds = data_dict["seed_inds"].long()
    seed_gt_votes_mask = torch.gather(data_dict["vote_label_mask"], 1, seed_inds)
    seed_inds_expand = seed_inds.view(batch_size, num_seed, 1).expand(-1, -1, 3 * GT_VOTE_FACTOR)
    seed_gt_votes = torch.gather(data_dict["vote_label"], 1, seed_inds_expand)
    seed_gt_votes = seed_gt_votes + data_dict["seed_xyz"].repeat(1, 1, 3)
    vote_xyz_reshape = vote_xyz.view(batch_size * num_seed, -1, 3)
    seed_gt_votes_reshape = seed_gt_votes.view(batch_size * num_seed, GT_VOTE_FACTOR, 3)
    _, _, dist2, _ = nn_distance(vote_xyz_reshape, seed_gt_votes_reshape, l1=True)
    votes_dist, _ = torch.min(dist2, dim=1)
    votes_dist = votes_dist.view(batch_size, num_seed)
    m = seed_gt_votes_mask.float()
    return torch.sum(votes_dist * m) / (torch.sum(m) + 1e-6)


def compute_objectness_loss(data_dict):
    aggregated_vote_xyz = data_dict["aggregated_vote_xyz"]
    gt_center = data_dict["center_label"][:, :, 0:3]
    dist1, ind1, _, _ = nn_distance(aggregated_vote_xyz, gt_center)
    euclidean_dist1 = torch.sqrt(dist1 + 1e-6)
    near = euclidean_dist1 < NEAR_THRESHOLD
    objectness_label = near.long()
    objectness_mask = (near | (euclidean_dist1 > FAR_THRESHOLD)).float()
    objectness_scores = data_dict["objectness_scores"]
    w = _const("obj_w", lambda: torch.tensor(OBJECTNESS_CLS_WEIGHTS, dtype=torch.float32), objectness_scores.device)
    objectness_loss = F.cross_entropy(objectness_scores.transpose(2, 1), objectness_label,
                                      weight=w.to(objectness_scores.dtype), reduction="none")
    objectness_loss = torch.sum(objectness_loss * objectness_mask) / (torch.sum(objectness_mask) + 1e-6)
    return objectness_loss, objectness_label, objectness_mask, ind1


def compute_box_and_sem_cls_loss(data_dict, config):
    num_heading_bin = config.num_heading_bin
    num_size_cluster = config.num_size_cluster
    mean_size_arr = config.mean_size_arr
    object_assignment = data_dict["object_assignment"]

    pred_center = data_dict["center"]
    gt_center = data_dict["center_label"][:, :, 0:3]
    dist1, _, dist2, _ = nn_distance(pred_center, gt_center)
    box_label_mask = data_dict["box_label_mask"]
    objectness_label = data_dict["objectness_label"].float()
    denom = torch.sum(objectness_label) + 1e-6
    centroid_reg_loss1 = torch.sum(dist1 * objectness_label) / denom
    centroid_reg_loss2 = torch.sum(dist2 * box_label_mask) / (torch.sum(box_label_mask) + 1e-6)
    center_loss = centroid_reg_loss1 + centroid_reg_loss2

    heading_class_label = torch.gather(data_dict["heading_class_label"], 1, object_assignment)
    heading_class_loss = F.cross_entropy(data_dict["heading_scores"].transpose(2, 1), heading_class_label, reduction="none")
    heading_class_loss = torch.sum(heading_class_loss * objectness_label) / denom
    heading_residual_label = torch.gather(data_dict["heading_residual_label"], 1, object_assignment)
    heading_residual_normalized_label = heading_residual_label / (np.pi / num_heading_bin)
    heading_label_one_hot = F.one_hot(heading_class_label, num_heading_bin).to(pred_center.dtype)
    heading_residual_normalized_loss = huber_loss(
        torch.sum(data_dict["heading_residuals_normalized"] * heading_label_one_hot, -1) - heading_residual_normalized_label,
        delta=1.0)
    heading_residual_normalized_loss = torch.sum(heading_residual_normalized_loss * objectness_label) / denom

    size_class_label = torch.gather(data_dict["size_class_label"], 1, object_assignment)
    size_class_loss = F.cross_entropy(data_dict["size_scores"].transpose(2, 1), size_class_label, reduction="none")
    size_class_loss = torch.sum(size_class_loss * objectness_label) / denom
    size_residual_label = torch.gather(data_dict["size_residual_label"], 1, object_assignment.unsqueeze(-1).expand(-1, -1, 3))
    size_label_one_hot_tiled = F.one_hot(size_class_label, num_size_cluster).to(pred_center.dtype).unsqueeze(-1)
    predicted_size_residual_normalized = torch.sum(data_dict["size_residuals_normalized"] * size_label_one_hot_tiled, 2)
    mean_size_arr_expanded = _const(("mean_size", np.asarray(mean_size_arr, np.float32).tobytes()),
                                    lambda: torch.from_numpy(np.asarray(mean_size_arr, np.float32)),
                                    pred_center.device).to(pred_center.dtype).unsqueeze(0).unsqueeze(0)
    mean_size_label = torch.sum(size_label_one_hot_tiled * mean_size_arr_expanded, 2)
    size_residual_label_normalized = size_residual_label / mean_size_label
    size_residual_normalized_loss = torch.mean(
        huber_loss(predicted_size_residual_normalized - size_residual_label_normalized, delta=1.0), -1)
    size_residual_normalized_loss = torch.sum(size_residual_normalized_loss * objectness_label) / denom

    sem_cls_label = torch.gather(data_dict["sem_cls_label"], 1, object_assignment)
    sem_cls_loss = F.cross_entropy(data_dict["sem_cls_scores"].transpose(2, 1), sem_cls_label, reduction="none")
    sem_cls_loss = torch.sum(sem_cls_loss * objectness_label) / denom
    return (center_loss, heading_class_loss, heading_residual_normalized_loss, size_class_loss,
            size_residual_normalized_loss, sem_cls_loss)


def compute_cap_loss(data_dict, config, weights):
    pred_caps = data_dict["lang_cap"]  # (B,T,V)
    B, T, num_vocabs = pred_caps.shape
    target_caps = data_dict["lang_ids"][:, 1:T + 1]  # == [:, 1:num_words]
    cap_loss = F.cross_entropy(pred_caps.reshape(-1, num_vocabs), target_caps.reshape(-1), ignore_index=0,
                               reduction="none")
    good = data_dict["good_bbox_masks"]
    good_rep = good.unsqueeze(1).expand(B, T).reshape(-1).to(cap_loss.dtype)
    # The reference's denominator sum(good_bbox_masks repeated num_words-1 times) (:213-215) counts the teacher-forced
    # steps of ITS run, num_words = lang_len.max().  The engine may run more steps than that (T padded to a bucket so
    # that one captured graph serves many caption lengths; the extra positions have pad targets = zero loss and zero
    # gradient), so the count is taken from lang_len on the device -- identical (exact integers) when T is not padded.
    steps = (data_dict["lang_len"].max() - 1).clamp(min=1, max=T).to(cap_loss.dtype)
    cap_loss = torch.sum(cap_loss * good_rep) / (torch.sum(good.to(cap_loss.dtype)) * steps + 1e-6)
    # accuracy over the non-pad tokens of the good boxes (0 if there is no good box) -- masked, no indexing
    with torch.no_grad():
        tok = (target_caps != 0) & good.unsqueeze(1)
        hit = (pred_caps.argmax(-1) == target_caps) & tok
        ntok = tok.sum().float()
        cap_acc = torch.where(good.any(), hit.sum().float() / ntok, torch.zeros_like(ntok))
    return cap_loss, cap_acc


def radian_to_label(radians, num_bins=6):
    boundaries = _const("bins%d" % num_bins, lambda: torch.arange(np.pi / num_bins, np.pi - 1e-8, np.pi / num_bins),
                        radians.device)
    return torch.bucketize(radians, boundaries)


def _edge_slots(data_dict):
    """Compact source / target ids (B,E) of the stored edges and the mask of the first
    num_edge_source*num_edge_target slots of each scene (what the reference's per-scene slices keep)."""
    edge_indices = data_dict["edge_index"]
    n = (data_dict["num_edge_source"] * data_dict["num_edge_target"]).unsqueeze(1)
    E = edge_indices.shape[2]
    keep = torch.arange(E, device=edge_indices.device).unsqueeze(0) < n
    return edge_indices[:, 0].long(), edge_indices[:, 1].long(), keep


def compute_node_orientation_loss(data_dict, num_bins=6):
    object_assignment = data_dict["object_assignment"]
    edge_preds = data_dict["edge_orientations"]  # (B,E,num_bins)
    B, K = object_assignment.shape
    rot = torch.gather(data_dict["scene_object_rotations"], 1, object_assignment.view(B, K, 1, 1).expand(-1, -1, 3, 3))
    rot_masks = torch.gather(data_dict["scene_object_rotation_masks"], 1, object_assignment)
    src, tar, keep = _edge_slots(data_dict)
    E = src.shape[1]
    source_rot = torch.gather(rot, 1, src.view(B, E, 1, 1).expand(-1, -1, 3, 3))
    target_rot = torch.gather(rot, 1, tar.view(B, E, 1, 1).expand(-1, -1, 3, 3))
    # trace(R_s R_t^T) = sum_ij R_s[i,j] R_t[i,j]: the reference forms the 3x3 products with a batched GEMM and sums the
    # diagonal (loss_helper.py:286-287); one elementwise product + reduction gives the same trace without 20 480 tiny GEMMs
    trace = (source_rot * target_rot).sum((-1, -2))
    relative_rot = torch.acos(torch.clamp(0.5 * (trace - 1), -1, 1))
    labels = radian_to_label(relative_rot, num_bins)
    masks = (torch.gather(rot_masks, 1, src) * torch.gather(rot_masks, 1, tar)) * keep.to(rot_masks.dtype)
    loss = F.cross_entropy(edge_preds.reshape(B * E, -1), labels.reshape(-1), reduction="none")
    masks_f = masks.reshape(-1).to(loss.dtype)
    loss = (loss * masks_f).sum() / (masks_f.sum() + 1e-8)
    with torch.no_grad():
        hit = (edge_preds.argmax(-1).reshape(-1) == labels.reshape(-1)) & (masks.reshape(-1) == 1)
        acc = hit.sum().float() / (masks_f.sum().float() + 1e-8)
    return loss, acc


def compute_node_distance_loss(data_dict):
    gt_center = data_dict["center_label"][:, :, 0:3]
    object_assignment = data_dict["object_assignment"]
    gt_center = torch.gather(gt_center, 1, object_assignment.unsqueeze(-1).expand(-1, -1, 3))
    edge_preds = data_dict["edge_distances"]
    src, tar, keep = _edge_slots(data_dict)
    sc = torch.gather(gt_center, 1, src.unsqueeze(-1).expand(-1, -1, 3))
    tc = torch.gather(gt_center, 1, tar.unsqueeze(-1).expand(-1, -1, 3))
    labels = torch.norm(sc - tc, dim=2)
    k = keep.to(edge_preds.dtype)
    return (((edge_preds - labels) ** 2) * k).sum() / k.sum()  # nn.MSELoss over the kept edges


def _detection_terms_fused(data_dict, config, detection):
    """The detection part of get_scene_cap_loss on the fused kernel (csrc/loss.cu): one launch for all terms, labels
    and gradients.  Returns 10 * (vote + 0.5 objectness + box + 0.1 sem_cls) or None when detection is off."""
    det, terms, objectness_label, objectness_mask, object_assignment = fused_loss.detection_loss(data_dict, config)
    data_dict["objectness_label"] = objectness_label
    data_dict["objectness_mask"] = objectness_mask
    data_dict["object_assignment"] = object_assignment
    data_dict["pos_ratio"], data_dict["neg_ratio"], data_dict["obj_acc"] = terms["pos_ratio"], terms["neg_ratio"], terms["obj_acc"]
    zero = torch.zeros((), device=det.device)
    for n in ("vote_loss", "objectness_loss", "center_loss", "heading_cls_loss", "heading_reg_loss", "size_cls_loss",
              "size_reg_loss", "sem_cls_loss", "box_loss"):
        data_dict[n] = terms[n] if detection else zero
    return det * 10 if detection else None


def get_scene_cap_loss(data_dict, device, config, weights, detection=True, caption=True, orientation=False,
                       distance=False, num_bins=CONF.TRAIN.NUM_BINS):
    if fused_loss.available(data_dict):
        det10 = _detection_terms_fused(data_dict, config, detection)
        return _finish_scene_cap_loss(data_dict, device, config, weights, det10, detection, caption, orientation,
                                      distance, num_bins)
    vote_loss = compute_vote_loss(data_dict)
    objectness_loss, objectness_label, objectness_mask, object_assignment = compute_objectness_loss(data_dict)
    total_num_proposal = objectness_label.shape[0] * objectness_label.shape[1]
    data_dict["objectness_label"] = objectness_label
    data_dict["objectness_mask"] = objectness_mask
    data_dict["object_assignment"] = object_assignment
    data_dict["pos_ratio"] = torch.sum(objectness_label.float()) / float(total_num_proposal)
    data_dict["neg_ratio"] = torch.sum(objectness_mask.float()) / float(total_num_proposal) - data_dict["pos_ratio"]

    center_loss, heading_cls_loss, heading_reg_loss, size_cls_loss, size_reg_loss, sem_cls_loss = \
        compute_box_and_sem_cls_loss(data_dict, config)
    box_loss = center_loss + 0.1 * heading_cls_loss + heading_reg_loss + 0.1 * size_cls_loss + size_reg_loss

    obj_pred_val = torch.argmax(data_dict["objectness_scores"], 2)
    data_dict["obj_acc"] = torch.sum((obj_pred_val == objectness_label.long()).float() * objectness_mask) / (
        torch.sum(objectness_mask) + 1e-6)

    zero = torch.zeros((), device=device)
    names = ["vote_loss", "objectness_loss", "center_loss", "heading_cls_loss", "heading_reg_loss", "size_cls_loss",
             "size_reg_loss", "sem_cls_loss", "box_loss"]
    vals = [vote_loss, objectness_loss, center_loss, heading_cls_loss, heading_reg_loss, size_cls_loss,
            size_reg_loss, sem_cls_loss, box_loss]
    for n, v in zip(names, vals):
        data_dict[n] = v if detection else zero

    det10 = None
    if detection:
        det10 = (data_dict["vote_loss"] + 0.5 * data_dict["objectness_loss"] + data_dict["box_loss"]
                 + 0.1 * data_dict["sem_cls_loss"]) * 10
    return _finish_scene_cap_loss(data_dict, device, config, weights, det10, detection, caption, orientation, distance,
                                  num_bins)


def _finish_scene_cap_loss(data_dict, device, config, weights, det10, detection, caption, orientation, distance, num_bins):
    """Caption / orientation / distance terms and the total (lib/loss_helper.py:436-491); det10 = 10 x detection loss."""
    zero = torch.zeros((), device=device)
    if caption:
        data_dict["cap_loss"], data_dict["cap_acc"] = compute_cap_loss(data_dict, config, weights)
    else:
        data_dict["cap_loss"], data_dict["cap_acc"], data_dict["pred_ious"] = zero, zero, zero
    if orientation:
        data_dict["ori_loss"], data_dict["ori_acc"] = compute_node_orientation_loss(data_dict, num_bins)
    else:
        data_dict["ori_loss"], data_dict["ori_acc"] = zero, zero
    data_dict["dist_loss"] = compute_node_distance_loss(data_dict) if distance else zero

    if detection:
        loss = det10
        if caption:
            loss = loss + data_dict["cap_loss"]
    else:
        loss = data_dict["cap_loss"]
    if orientation:
        loss = loss + 0.1 * data_dict["ori_loss"]
    if distance:
        loss = loss + 0.1 * data_dict["dist_loss"]
    data_dict["loss"] = loss
    return data_dict
