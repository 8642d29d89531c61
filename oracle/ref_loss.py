"""TEST INFRASTRUCTURE ONLY -- literal restatement of the reference's loss (lib/loss_helper.py:24-491,
utils/nn_distance.py:11-59) with its per-scene Python loops, boolean-mask indexing and host-side branches
kept as they are; ``.cuda()`` became ``.to(device)``.  Used to check scan2cap_b200/lib/loss_helper.py (the
masked, synchronisation-free version) and inside the reference arm of bench.py."""
import numpy as np
import torch
import torch.nn as nn

FAR_THRESHOLD = 0.6
NEAR_THRESHOLD = 0.3
GT_VOTE_FACTOR = 3
OBJECTNESS_CLS_WEIGHTS = [0.2, 0.8]


def huber_loss(error, delta=1.0):  # nn_distance.py:11-28
    abs_error = torch.abs(error)
    quadratic = torch.clamp(abs_error, max=delta)
    linear = (abs_error - quadratic)
    return 0.5 * quadratic ** 2 + delta * linear


def nn_distance(pc1, pc2, l1smooth=False, delta=1.0, l1=False):  # nn_distance.py:32-59
    N, M = pc1.shape[1], pc2.shape[1]
    pc_diff = pc1.unsqueeze(2).repeat(1, 1, M, 1) - pc2.unsqueeze(1).repeat(1, N, 1, 1)
    if l1smooth:
        pc_dist = torch.sum(huber_loss(pc_diff, delta), dim=-1)
    elif l1:
        pc_dist = torch.sum(torch.abs(pc_diff), dim=-1)
    else:
        pc_dist = torch.sum(pc_diff ** 2, dim=-1)
    dist1, idx1 = torch.min(pc_dist, dim=2)
    dist2, idx2 = torch.min(pc_dist, dim=1)
    return dist1, idx1, dist2, idx2


def compute_vote_loss(d):  # loss_helper.py:24-69
    batch_size, num_seed = d["seed_xyz"].shape[0], d["seed_xyz"].shape[1]
    vote_xyz = d["vote_xyz"]
    seed_inds = d["seed_inds"].long()
    seed_gt_votes_mask = torch.gather(d["vote_label_mask"], 1, seed_inds)
    seed_inds_expand = seed_inds.view(batch_size, num_seed, 1).repeat(1, 1, 3 * GT_VOTE_FACTOR)
    seed_gt_votes = torch.gather(d["vote_label"], 1, seed_inds_expand)
    seed_gt_votes += d["seed_xyz"].repeat(1, 1, 3)
    vote_xyz_reshape = vote_xyz.view(batch_size * num_seed, -1, 3)
    seed_gt_votes_reshape = seed_gt_votes.view(batch_size * num_seed, GT_VOTE_FACTOR, 3)
    dist1, _, dist2, _ = nn_distance(vote_xyz_reshape, seed_gt_votes_reshape, l1=True)
    votes_dist, _ = torch.min(dist2, dim=1)
    votes_dist = votes_dist.view(batch_size, num_seed)
    return torch.sum(votes_dist * seed_gt_votes_mask.float()) / (torch.sum(seed_gt_votes_mask.float()) + 1e-6)


def compute_objectness_loss(d):  # :71-111
    aggregated_vote_xyz = d["aggregated_vote_xyz"]
    dev = aggregated_vote_xyz.device
    gt_center = d["center_label"][:, :, 0:3]
    B, K = gt_center.shape[0], aggregated_vote_xyz.shape[1]
    dist1, ind1, dist2, _ = nn_distance(aggregated_vote_xyz, gt_center)
    euclidean_dist1 = torch.sqrt(dist1 + 1e-6)
    objectness_label = torch.zeros((B, K), dtype=torch.long, device=dev)
    objectness_mask = torch.zeros((B, K), device=dev)
    objectness_label[euclidean_dist1 < NEAR_THRESHOLD] = 1
    objectness_mask[euclidean_dist1 < NEAR_THRESHOLD] = 1
    objectness_mask[euclidean_dist1 > FAR_THRESHOLD] = 1
    criterion = nn.CrossEntropyLoss(torch.Tensor(OBJECTNESS_CLS_WEIGHTS).to(dev), reduction="none")
    objectness_loss = criterion(d["objectness_scores"].transpose(2, 1), objectness_label)
    objectness_loss = torch.sum(objectness_loss * objectness_mask) / (torch.sum(objectness_mask) + 1e-6)
    return objectness_loss, objectness_label, objectness_mask, ind1


def compute_box_and_sem_cls_loss(d, config):  # :113-187
    num_heading_bin, num_size_cluster, mean_size_arr = config.num_heading_bin, config.num_size_cluster, config.mean_size_arr
    object_assignment = d["object_assignment"]
    batch_size = object_assignment.shape[0]
    dev = object_assignment.device
    pred_center = d["center"]
    gt_center = d["center_label"][:, :, 0:3]
    dist1, ind1, dist2, _ = nn_distance(pred_center, gt_center)
    box_label_mask = d["box_label_mask"]
    objectness_label = d["objectness_label"].float()
    centroid_reg_loss1 = torch.sum(dist1 * objectness_label) / (torch.sum(objectness_label) + 1e-6)
    centroid_reg_loss2 = torch.sum(dist2 * box_label_mask) / (torch.sum(box_label_mask) + 1e-6)
    center_loss = centroid_reg_loss1 + centroid_reg_loss2
    heading_class_label = torch.gather(d["heading_class_label"], 1, object_assignment)
    heading_class_loss = nn.CrossEntropyLoss(reduction="none")(d["heading_scores"].transpose(2, 1), heading_class_label)
    heading_class_loss = torch.sum(heading_class_loss * objectness_label) / (torch.sum(objectness_label) + 1e-6)
    heading_residual_label = torch.gather(d["heading_residual_label"], 1, object_assignment)
    heading_residual_normalized_label = heading_residual_label / (np.pi / num_heading_bin)
    heading_label_one_hot = torch.zeros(batch_size, heading_class_label.shape[1], num_heading_bin, device=dev)
    heading_label_one_hot.scatter_(2, heading_class_label.unsqueeze(-1), 1)
    heading_residual_normalized_loss = huber_loss(
        torch.sum(d["heading_residuals_normalized"] * heading_label_one_hot, -1) - heading_residual_normalized_label, delta=1.0)
    heading_residual_normalized_loss = torch.sum(heading_residual_normalized_loss * objectness_label) / (torch.sum(objectness_label) + 1e-6)
    size_class_label = torch.gather(d["size_class_label"], 1, object_assignment)
    size_class_loss = nn.CrossEntropyLoss(reduction="none")(d["size_scores"].transpose(2, 1), size_class_label)
    size_class_loss = torch.sum(size_class_loss * objectness_label) / (torch.sum(objectness_label) + 1e-6)
    size_residual_label = torch.gather(d["size_residual_label"], 1, object_assignment.unsqueeze(-1).repeat(1, 1, 3))
    size_label_one_hot = torch.zeros(batch_size, size_class_label.shape[1], num_size_cluster, device=dev)
    size_label_one_hot.scatter_(2, size_class_label.unsqueeze(-1), 1)
    size_label_one_hot_tiled = size_label_one_hot.unsqueeze(-1).repeat(1, 1, 1, 3)
    predicted_size_residual_normalized = torch.sum(d["size_residuals_normalized"] * size_label_one_hot_tiled, 2)
    mean_size_arr_expanded = torch.from_numpy(mean_size_arr.astype(np.float32)).to(dev).unsqueeze(0).unsqueeze(0)
    mean_size_label = torch.sum(size_label_one_hot_tiled * mean_size_arr_expanded, 2)
    size_residual_label_normalized = size_residual_label / mean_size_label
    size_residual_normalized_loss = torch.mean(huber_loss(predicted_size_residual_normalized - size_residual_label_normalized, delta=1.0), -1)
    size_residual_normalized_loss = torch.sum(size_residual_normalized_loss * objectness_label) / (torch.sum(objectness_label) + 1e-6)
    sem_cls_label = torch.gather(d["sem_cls_label"], 1, object_assignment)
    sem_cls_loss = nn.CrossEntropyLoss(reduction="none")(d["sem_cls_scores"].transpose(2, 1), sem_cls_label)
    sem_cls_loss = torch.sum(sem_cls_loss * objectness_label) / (torch.sum(objectness_label) + 1e-6)
    return center_loss, heading_class_loss, heading_residual_normalized_loss, size_class_loss, size_residual_normalized_loss, sem_cls_loss


def compute_cap_loss(d):  # :189-230
    pred_caps = d["lang_cap"]
    dev = pred_caps.device
    num_words = d["lang_len"].max()
    target_caps = d["lang_ids"][:, 1:num_words]
    _, _, num_vocabs = pred_caps.shape
    cap_loss = nn.CrossEntropyLoss(ignore_index=0, reduction="none")(pred_caps.reshape(-1, num_vocabs), target_caps.reshape(-1))
    good_bbox_masks = d["good_bbox_masks"].unsqueeze(1).repeat(1, num_words - 1).reshape(-1)
    cap_loss = torch.sum(cap_loss * good_bbox_masks) / (torch.sum(good_bbox_masks) + 1e-6)
    num_good_bbox = d["good_bbox_masks"].sum()
    if num_good_bbox > 0:
        pc = pred_caps[d["good_bbox_masks"]].reshape(-1, num_vocabs).argmax(-1)
        tc = target_caps[d["good_bbox_masks"]].reshape(-1)
        masks = tc != 0
        cap_acc = (pc[masks] == tc[masks]).sum().float() / masks.sum().float()
    else:
        cap_acc = torch.zeros(1, device=dev)[0]
    return cap_loss, cap_acc


def radian_to_label(radians, num_bins=6):  # :232-248
    boundaries = torch.arange(np.pi / num_bins, np.pi - 1e-8, np.pi / num_bins).to(radians.device)
    return torch.bucketize(radians, boundaries)


def compute_node_orientation_loss(d, num_bins=6):  # :250-313
    object_assignment = d["object_assignment"]
    edge_indices, edge_preds = d["edge_index"], d["edge_orientations"]
    num_sources, num_targets = d["num_edge_source"], d["num_edge_target"]
    batch_size, num_proposals = object_assignment.shape
    object_rotation_matrices = torch.gather(d["scene_object_rotations"], 1,
                                            object_assignment.view(batch_size, num_proposals, 1, 1).repeat(1, 1, 3, 3))
    object_rotation_masks = torch.gather(d["scene_object_rotation_masks"], 1, object_assignment)
    preds, labels, masks = [], [], []
    for batch_id in range(batch_size):
        batch_rotations = object_rotation_matrices[batch_id]
        batch_rotation_masks = object_rotation_masks[batch_id]
        n = num_sources[batch_id] * num_targets[batch_id]
        source_indices = edge_indices[batch_id, 0, :n].long()
        target_indices = edge_indices[batch_id, 1, :n].long()
        source_rot = torch.index_select(batch_rotations, 0, source_indices)
        target_rot = torch.index_select(batch_rotations, 0, target_indices)
        relative_rot = torch.matmul(source_rot, target_rot.transpose(2, 1))
        relative_rot = torch.acos(torch.clamp(0.5 * (torch.diagonal(relative_rot, dim1=-2, dim2=-1).sum(-1) - 1), -1, 1))
        source_masks = torch.index_select(batch_rotation_masks, 0, source_indices)
        target_masks = torch.index_select(batch_rotation_masks, 0, target_indices)
        preds.append(edge_preds[batch_id, :n])
        labels.append(radian_to_label(relative_rot, num_bins))
        masks.append(source_masks * target_masks)
    preds, labels, masks = torch.cat(preds, dim=0), torch.cat(labels, dim=0), torch.cat(masks, dim=0)
    loss = nn.CrossEntropyLoss(reduction="none")(preds, labels)
    loss = (loss * masks).sum() / (masks.sum() + 1e-8)
    preds = preds.argmax(-1)
    acc = (preds[masks == 1] == labels[masks == 1]).sum().float() / (masks.sum().float() + 1e-8)
    return loss, acc


def compute_node_distance_loss(d):  # :315-355
    gt_center = d["center_label"][:, :, 0:3]
    object_assignment = d["object_assignment"]
    gt_center = torch.gather(gt_center, 1, object_assignment.unsqueeze(-1).repeat(1, 1, 3))
    batch_size = gt_center.shape[0]
    preds, labels = [], []
    for batch_id in range(batch_size):
        n = d["num_edge_source"][batch_id] * d["num_edge_target"][batch_id]
        source_indices = d["edge_index"][batch_id, 0, :n].long()
        target_indices = d["edge_index"][batch_id, 1, :n].long()
        sc = torch.index_select(gt_center[batch_id], 0, source_indices)
        tc = torch.index_select(gt_center[batch_id], 0, target_indices)
        labels.append(torch.norm(sc - tc, dim=1))
        preds.append(d["edge_distances"][batch_id, :n])
    return nn.MSELoss()(torch.cat(preds, dim=0), torch.cat(labels, dim=0))


def get_scene_cap_loss(d, device, config, weights=None, detection=True, caption=True, orientation=False,
                       distance=False, num_bins=6):  # :381-491
    vote_loss = compute_vote_loss(d)
    objectness_loss, objectness_label, objectness_mask, object_assignment = compute_objectness_loss(d)
    total_num_proposal = objectness_label.shape[0] * objectness_label.shape[1]
    d["objectness_label"], d["objectness_mask"], d["object_assignment"] = objectness_label, objectness_mask, object_assignment
    d["pos_ratio"] = torch.sum(objectness_label.float().to(device)) / float(total_num_proposal)
    d["neg_ratio"] = torch.sum(objectness_mask.float()) / float(total_num_proposal) - d["pos_ratio"]
    center_loss, heading_cls_loss, heading_reg_loss, size_cls_loss, size_reg_loss, sem_cls_loss = compute_box_and_sem_cls_loss(d, config)
    box_loss = center_loss + 0.1 * heading_cls_loss + heading_reg_loss + 0.1 * size_cls_loss + size_reg_loss
    obj_pred_val = torch.argmax(d["objectness_scores"], 2)
    d["obj_acc"] = torch.sum((obj_pred_val == d["objectness_label"].long()).float() * d["objectness_mask"]) / (torch.sum(d["objectness_mask"]) + 1e-6)
    z = lambda: torch.zeros(1)[0].to(device)
    if detection:
        d["vote_loss"], d["objectness_loss"], d["center_loss"] = vote_loss, objectness_loss, center_loss
        d["heading_cls_loss"], d["heading_reg_loss"] = heading_cls_loss, heading_reg_loss
        d["size_cls_loss"], d["size_reg_loss"], d["sem_cls_loss"], d["box_loss"] = size_cls_loss, size_reg_loss, sem_cls_loss, box_loss
    else:
        for k in ("vote_loss", "objectness_loss", "center_loss", "heading_cls_loss", "heading_reg_loss", "size_cls_loss",
                  "size_reg_loss", "sem_cls_loss", "box_loss"):
            d[k] = z()
    if caption:
        d["cap_loss"], d["cap_acc"] = compute_cap_loss(d)
    else:
        d["cap_loss"], d["cap_acc"], d["pred_ious"] = z(), z(), z()
    if orientation:
        d["ori_loss"], d["ori_acc"] = compute_node_orientation_loss(d, num_bins)
    else:
        d["ori_loss"], d["ori_acc"] = z(), z()
    d["dist_loss"] = compute_node_distance_loss(d) if distance else z()
    if detection:
        loss = d["vote_loss"] + 0.5 * d["objectness_loss"] + d["box_loss"] + 0.1 * d["sem_cls_loss"]
        loss *= 10
        if caption:
            loss += d["cap_loss"]
        if orientation:
            loss += 0.1 * d["ori_loss"]
        if distance:
            loss += 0.1 * d["dist_loss"]
    else:
        loss = d["cap_loss"]
        if orientation:
            loss += 0.1 * d["ori_loss"]
        if distance:
            loss += 0.1 * d["dist_loss"]
    d["loss"] = loss
    return d
