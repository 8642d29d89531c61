"""Mirror of lib/loss_helper_pretrained.py (get_loss :167-214): the loss of the capnet_pretrained path
(BASELINE configs[0]: graph + caption on pre-extracted box features) -- caption cross-entropy (:16-78) and the
relative-orientation loss of the graph edges (:98-165, with object_assignment taken from bbox_center vs
bbox_center_label).  Sync-free like lib/loss_helper.py: masked reductions instead of boolean indexing / per-scene
loops, no .item()."""
import torch
import torch.nn.functional as F

from ..utils.nn_distance import nn_distance
from .config import CONF
from .loss_helper import compute_node_orientation_loss as _orientation_loss


def compute_cap_loss(data_dict, mode="gt"):
    pred_caps = data_dict["lang_cap"]  # (B,T,V)
    B, T, num_vocabs = pred_caps.shape
    target_caps = data_dict["lang_ids"][:, 1:T + 1]  # == [:, 1:num_words]
    with torch.no_grad():
        hit = pred_caps.argmax(-1) == target_caps
        tok = target_caps != 0
    if mode == "gt":
        # nn.CrossEntropyLoss(ignore_index=0), mean over the non-pad tokens (:31-32)
        ce = F.cross_entropy(pred_caps.reshape(-1, num_vocabs), target_caps.reshape(-1), ignore_index=0, reduction="none")
        cap_loss = ce.sum() / tok.sum().to(ce.dtype)
        cap_acc = (hit & tok).sum().float() / tok.sum().float()
        return cap_loss, cap_acc
    ce = F.cross_entropy(pred_caps.reshape(-1, num_vocabs), target_caps.reshape(-1), ignore_index=0, reduction="none")
    good = data_dict["good_bbox_masks"]
    good_rep = good.unsqueeze(1).expand(B, T).reshape(-1).to(ce.dtype)
    # denominator = good boxes x teacher-forced steps of the reference's run (see lib/loss_helper.py::compute_cap_loss)
    steps = (data_dict["lang_len"].max() - 1).clamp(min=1, max=T).to(ce.dtype)
    cap_loss = torch.sum(ce * good_rep) / (torch.sum(good.to(ce.dtype)) * steps + 1e-6)
    with torch.no_grad():
        tok_g = tok & good.unsqueeze(1)
        ntok = tok_g.sum().float()
        cap_acc = torch.where(good.any(), (hit & tok_g).sum().float() / ntok, torch.zeros_like(ntok))
    return cap_loss, cap_acc


def compute_node_orientation_loss(data_dict, num_bins=6):
    _, object_assignment, _, _ = nn_distance(data_dict["bbox_center"], data_dict["bbox_center_label"])  # f32 - f64 -> f64
    d = dict(data_dict)
    d["object_assignment"] = object_assignment
    return _orientation_loss(d, num_bins)


def get_loss(data_dict, mode="gt", orientation=False, num_bins=CONF.TRAIN.NUM_BINS):
    data_dict["cap_loss"], data_dict["cap_acc"] = compute_cap_loss(data_dict, mode)
    if orientation:
        data_dict["ori_loss"], data_dict["ori_acc"] = compute_node_orientation_loss(data_dict, num_bins)
    else:
        zero = torch.zeros((), device=data_dict["lang_cap"].device)
        data_dict["ori_loss"], data_dict["ori_acc"] = zero, zero
    loss = data_dict["cap_loss"]
    if orientation:
        # the reference accumulates IN PLACE (`loss = data_dict["cap_loss"]; loss += 0.1 * ori_loss`, :207-209), so the
        # "cap_loss" it logs afterwards is the total loss; reproduced
        loss = loss + 0.1 * data_dict["ori_loss"]
        data_dict["cap_loss"] = loss
    data_dict["loss"] = loss
    return data_dict
