"""Mirror of lib/ap_helper.py::parse_predictions (:40-178) -- the post-processing benchmark/predict.py applies to every
batch (:176-190): decode boxes, drop boxes with fewer than 5 scene points inside, class-aware 3-D NMS, per-class
prediction lists -- computed on the device.

The reference does this on the host: B*K python iterations that each build a box and a scipy Delaunay hull of it, then
a numpy NMS loop per scene; after the model itself is fast that loop dominates predict.py's wall time (SURVEY 8(f) row
2).  Here: the box decode is the float64 tensor arithmetic of ProposalModule.decode_pred_box (ScanNet boxes have
heading 0, model_util_scannet.py:130-134, so a box is its min / max corner), the point count and the NMS are two
libs2c launches (csrc/nms.cu), and ONE device->host transfer brings back what predict.py needs."""
import numpy as np
import torch

from .._lib import call
from .pointnet2._ext import _guard, _stream


def points_in_boxes_count(xyz, boxes):
    """xyz (B,N,>=3) fp32 with unit last stride (e.g. point_clouds), boxes (B,K,6) f64 [min xyz | max xyz] -> (B,K) int32."""
    assert xyz.is_cuda and xyz.dtype == torch.float32 and xyz.stride(2) == 1 and xyz.stride(0) == xyz.shape[1] * xyz.stride(1)
    boxes = boxes.contiguous()
    B, N, K = xyz.shape[0], xyz.shape[1], boxes.shape[1]
    out = torch.empty((B, K), dtype=torch.int32, device=xyz.device)
    with _guard(xyz):
        call("s2c_points_in_boxes_count", xyz.data_ptr(), xyz.stride(1), B, N, boxes.data_ptr(), K, out.data_ptr(),
             _stream(xyz))
    return out


def nms3d(boxes, score, cls, valid, iou_threshold, old_type=False, same_class_only=True):
    """boxes (B,K,6) f64, score (B,K) f64, cls (B,K) int64, valid (B,K) int32 -> keep (B,K) int32 (utils/nms.py)."""
    boxes, score, cls, valid = boxes.contiguous(), score.contiguous(), cls.contiguous(), valid.contiguous()
    B, K = score.shape
    keep = torch.empty((B, K), dtype=torch.int32, device=boxes.device)
    with _guard(boxes):
        call("s2c_nms3d", boxes.data_ptr(), score.data_ptr(), cls.data_ptr(), valid.data_ptr(), B, K, float(iou_threshold),
             1 if old_type else 0, 1 if same_class_only else 0, keep.data_ptr(), _stream(boxes))
    return keep


def parse_predictions_device(end_points, config_dict):
    """The device part: -> dict of device tensors: corners (B,K,8,3) f64, pred_mask (B,K) int32, nonempty (B,K) int32,
    obj_prob (B,K) f32, sem_cls_probs (B,K,C) f32, pred_sem_cls (B,K) int64."""
    assert config_dict["use_3d_nms"], "only the 3-D NMS variants (the predict.py / eval configuration) are provided"
    DC = config_dict["dataset_config"]
    center = end_points["center"].detach()
    dev = center.device
    size_class = torch.argmax(end_points["size_scores"], -1)
    size_residual = torch.gather(end_points["size_residuals"].detach(), 2,
                                 size_class.view(*size_class.shape, 1, 1).expand(-1, -1, 1, 3)).squeeze(2)
    mean = torch.as_tensor(np.asarray(DC.mean_size_arr, np.float64), device=dev)
    box_size = mean[size_class] + size_residual.double()          # class2size (model_util_scannet.py:148-150)
    c64 = center.double()                                        # (heading is 0: get_3d_box is centre +- size / 2)
    lo, hi = c64 - box_size / 2, c64 + box_size / 2
    # min / max over the corners, as the reference takes them (:104-111): a negative predicted size flips lo and hi
    boxes = torch.cat([torch.minimum(lo, hi), torch.maximum(lo, hi)], -1)   # (B,K,6)
    sx = torch.tensor([1, 1, -1, -1, 1, 1, -1, -1], dtype=torch.float64, device=dev)
    sy = torch.tensor([1, -1, -1, 1, 1, -1, -1, 1], dtype=torch.float64, device=dev)
    sz = torch.tensor([1, 1, 1, 1, -1, -1, -1, -1], dtype=torch.float64, device=dev)
    corners = c64.unsqueeze(2) + torch.stack([sx, sy, sz], -1) * (box_size.unsqueeze(2) / 2)   # box_util.py:340-358
    sem_scores = end_points["sem_cls_scores"].detach()
    pred_sem_cls = torch.argmax(sem_scores, -1)
    sem_cls_probs = torch.softmax(sem_scores, -1)
    obj_prob = torch.softmax(end_points["objectness_scores"].detach(), -1)[:, :, 1]
    if config_dict["remove_empty_box"]:
        nonempty = (points_in_boxes_count(end_points["point_clouds"], boxes) >= 5).int()
    else:
        nonempty = torch.ones(obj_prob.shape, dtype=torch.int32, device=dev)
    pred_mask = nms3d(boxes, obj_prob.double(), pred_sem_cls, nonempty, config_dict["nms_iou"],
                      config_dict["use_old_type_nms"], bool(config_dict["cls_nms"]))
    return {"corners": corners, "pred_mask": pred_mask, "nonempty": nonempty, "obj_prob": obj_prob,
            "sem_cls_probs": sem_cls_probs, "pred_sem_cls": pred_sem_cls}


def parse_predictions(end_points, config_dict):
    """Same contract as the reference: sets end_points["pred_mask"] ((B,K) numpy 0/1) and
    end_points["batch_pred_map_cls"], returns the latter (list over scenes of (class, corners (8,3), score))."""
    dev_out = parse_predictions_device(end_points, config_dict)
    host = {k: v.cpu().numpy() for k, v in dev_out.items()}      # the one transfer
    pred_mask, obj_prob, corners = host["pred_mask"].astype(np.float64), host["obj_prob"], host["corners"]
    end_points["pred_mask"] = pred_mask
    DC = config_dict["dataset_config"]
    out = []
    for i in range(pred_mask.shape[0]):
        sel = [j for j in range(pred_mask.shape[1]) if pred_mask[i, j] == 1 and obj_prob[i, j] > config_dict["conf_thresh"]]
        if config_dict["per_class_proposal"]:
            cur = []
            for ii in range(DC.num_class):
                cur += [(ii, corners[i, j], host["sem_cls_probs"][i, j, ii] * obj_prob[i, j]) for j in sel]
        else:
            cur = [(int(host["pred_sem_cls"][i, j]), corners[i, j], obj_prob[i, j]) for j in sel]
        out.append(cur)
    end_points["batch_pred_map_cls"] = out
    return out
