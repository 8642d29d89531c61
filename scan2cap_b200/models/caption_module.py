"""Mirror of models/caption_module.py: select_target (:16-38), SceneCaptionModule (:40-200) and
TopDownSceneCaptionModule (:202-592) -- same constructor arguments, attribute names (map_topdown,
recurrent_cell_1/2, map_feat, map_hidd, attend, map_lang, classifier -> checkpoint keys) and ``data_dict``
outputs (lang_cap, pred_ious, topdown_attn, valid_masks, good_bbox_masks).

Differences in how the work is issued (results are the same):
  * select_target is batched: one IoU evaluation for all scenes, arg-max on the device, no .item();
  * the k-nearest "local context" mask of the targets is one s2c_knn_adjacency launch;
  * map_feat(obj_feats) is loop-invariant and is hoisted out of the recurrence (the reference recomputes
    it at every step, caption_module.py:275);
  * evaluation (256 proposals x 29 greedy steps) runs all B*256 sequences together and looks the next
    embedding up in a device-side table instead of per-token .item() + dict lookup + H2D copy
    (caption_module.py:553-566).
The number of teacher-forced steps is data dependent (lang_len.max()); pass ``data_dict["num_words"]`` (a
Python int, e.g. computed by the loader) to avoid the one device->host read this otherwise needs.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..lib import caption_decoder
from ..lib.linear_simt import linear
from ..lib.config import CONF
from ..lib.pointnet2 import _ext_graph
from ..utils.box_util import box3d_iou_batch_tensor


def select_target(data_dict):
    """Proposal with the highest IoU against the referred GT box, per scene -> (ids (B) int64, ious (B) f32)."""
    pred_bbox = data_dict["bbox_corner"]  # (B,K,8,3)
    B, K = pred_bbox.shape[0], pred_bbox.shape[1]
    gt_bbox = data_dict["ref_box_corner_label"].to(pred_bbox.dtype)  # (B,8,3)
    ious = box3d_iou_batch_tensor(pred_bbox.reshape(B * K, 8, 3),
                                  gt_bbox.unsqueeze(1).expand(B, K, 8, 3).reshape(B * K, 8, 3)).view(B, K)
    target_ious, target_ids = ious.max(dim=1)
    return target_ids, target_ious.float()


def _num_words(data_dict):
    n = data_dict.get("num_words", None)
    if n is None:
        n = int(data_dict["lang_len"].max().item())
    return int(n)


class SceneCaptionModule(nn.Module):
    def __init__(self, vocabulary, embeddings, emb_size=300, feat_size=128, hidden_size=512, num_proposals=256):
        super().__init__()
        self.vocabulary = vocabulary
        self.embeddings = embeddings
        self.num_vocabs = len(vocabulary["word2idx"])
        self.emb_size = emb_size
        self.feat_size = feat_size
        self.hidden_size = hidden_size
        self.num_proposals = num_proposals
        self.map_feat = nn.Sequential(nn.Linear(feat_size, emb_size), nn.ReLU())
        self.recurrent_cell = nn.GRUCell(input_size=emb_size, hidden_size=emb_size)
        self.classifier = nn.Linear(emb_size, self.num_vocabs)
        self._emb_table = None

    def step(self, step_input, hidden):
        hidden = self.recurrent_cell(step_input, hidden)
        return hidden, hidden

    def forward(self, data_dict, use_tf=True, is_eval=False, max_len=CONF.TRAIN.MAX_DES_LEN):
        if not is_eval:
            return self.forward_sample_batch(data_dict, max_len)
        return self.forward_scene_batch(data_dict, use_tf, max_len)

    def forward_sample_batch(self, data_dict, max_len=CONF.TRAIN.MAX_DES_LEN, min_iou=CONF.TRAIN.MIN_IOU_THRESHOLD):
        word_embs = data_dict["lang_feat"]
        num_words = _num_words(data_dict)
        obj_feats = self.map_feat(data_dict["bbox_feature"])
        B = obj_feats.shape[0]
        target_ids, target_ious = select_target(data_dict)
        hidden = torch.gather(obj_feats, 1, target_ids.view(B, 1, 1).expand(B, 1, self.emb_size)).squeeze(1)
        outputs = []
        for step_id in range(max(num_words - 1, 1)):
            step_output, hidden = self.step(word_embs[:, step_id], hidden)
            outputs.append(self.classifier(step_output).unsqueeze(1))
        good = target_ious > min_iou
        data_dict["lang_cap"] = torch.cat(outputs, dim=1)
        data_dict["pred_ious"] = _masked_mean(target_ious, good)
        data_dict["good_bbox_masks"] = good
        return data_dict

    def forward_scene_batch(self, data_dict, use_tf=False, max_len=CONF.TRAIN.MAX_DES_LEN):
        word_embs = data_dict["lang_feat"]
        obj_feats = self.map_feat(data_dict["bbox_feature"])
        B, K, _ = obj_feats.shape
        steps = (_num_words(data_dict) - 1) if use_tf else (max_len - 1)
        table = _embedding_table(self, obj_feats.device)
        hidden = obj_feats.reshape(B * K, self.emb_size)
        step_input = word_embs[:, 0].unsqueeze(1).expand(B, K, self.emb_size).reshape(B * K, self.emb_size)
        outputs = []
        for step_id in range(steps):
            step_output, hidden = self.step(step_input, hidden)
            logits = self.classifier(step_output)
            outputs.append(logits.view(B, K, 1, -1))
            if use_tf:
                nxt = word_embs[:, min(step_id + 1, word_embs.shape[1] - 1)]
                step_input = nxt.unsqueeze(1).expand(B, K, self.emb_size).reshape(B * K, self.emb_size)
            else:
                step_input = table.index_select(0, logits.argmax(-1))
        data_dict["lang_cap"] = torch.cat(outputs, dim=2)
        return data_dict


def _masked_mean(values, mask):
    """values[mask].mean() if mask.any() else 0 -- without a host round trip."""
    m = mask.to(values.dtype)
    cnt = m.sum()
    return torch.where(cnt > 0, (values * m).sum() / cnt.clamp_min(1), torch.zeros_like(cnt))


def _embedding_table(module, device):
    """(V, emb) table with row i = embeddings[idx2word[str(i)]] (what the reference looks up per token)."""
    if module._emb_table is None or module._emb_table.device != device:
        idx2word = module.vocabulary["idx2word"]
        rows = [np.asarray(module.embeddings[idx2word[str(i)]], dtype=np.float32) for i in range(module.num_vocabs)]
        module._emb_table = torch.from_numpy(np.stack(rows)).to(device)
    return module._emb_table


class TopDownSceneCaptionModule(nn.Module):
    def __init__(self, vocabulary, embeddings, emb_size=300, feat_size=128, hidden_size=512, num_proposals=256,
                 num_locals=-1, query_mode="corner", use_relation=False, use_oracle=False):
        super().__init__()
        self.vocabulary = vocabulary
        self.embeddings = embeddings
        self.num_vocabs = len(vocabulary["word2idx"])
        self.emb_size = emb_size
        self.feat_size = feat_size
        self.hidden_size = hidden_size
        self.num_proposals = num_proposals
        self.num_locals = num_locals
        self.query_mode = query_mode
        self.use_relation = use_relation
        self.use_oracle = use_oracle
        self.map_topdown = nn.Sequential(nn.Linear(hidden_size + feat_size + emb_size, emb_size), nn.ReLU())
        self.recurrent_cell_1 = nn.GRUCell(input_size=emb_size, hidden_size=hidden_size)
        self.map_feat = nn.Linear(feat_size, hidden_size, bias=False)
        self.map_hidd = nn.Linear(hidden_size, hidden_size, bias=False)
        self.attend = nn.Linear(hidden_size, 1, bias=False)
        self.map_lang = nn.Sequential(nn.Linear(feat_size + hidden_size, emb_size), nn.ReLU())
        self.recurrent_cell_2 = nn.GRUCell(input_size=emb_size, hidden_size=hidden_size)
        self.classifier = nn.Linear(hidden_size, self.num_vocabs)
        self._emb_table = None

    # ---- one decoder step (caption_module.py:250-292); `mapped_feats` = map_feat(obj_feats), hoisted ----
    def _step(self, step_input, target_feat, obj_feats, hidden_1, hidden_2, object_masks, mapped_feats=None):
        step_input = self.map_topdown(torch.cat([step_input, hidden_2, target_feat], dim=-1))
        hidden_1 = self.recurrent_cell_1(step_input, hidden_1)
        if mapped_feats is None:
            mapped_feats = self.map_feat(obj_feats)
        combined = torch.tanh(mapped_feats + self.map_hidd(hidden_1).unsqueeze(1))
        scores = self.attend(combined).masked_fill(object_masks == 0, float("-1e30"))  # (B,K,1)
        masks = F.softmax(scores, dim=1)
        attended = (obj_feats * masks).sum(1)
        lang_input = self.map_lang(torch.cat([attended, hidden_1], dim=-1))
        hidden_2 = self.recurrent_cell_2(lang_input, hidden_2)
        return hidden_1, hidden_2, masks

    def _step_core(self, topdown_input, obj_feats, hidden_1, hidden_2, object_masks, mapped_feats):
        """_step after the map_topdown fusion layer (whose input-only terms the caller has hoisted)."""
        hidden_1 = self.recurrent_cell_1(topdown_input, hidden_1)
        combined = torch.tanh(mapped_feats + self.map_hidd(hidden_1).unsqueeze(1))
        scores = self.attend(combined).masked_fill(object_masks == 0, float("-1e30"))  # (B,K,1)
        masks = F.softmax(scores, dim=1)
        attended = (obj_feats * masks).sum(1)
        lang_input = self.map_lang(torch.cat([attended, hidden_1], dim=-1))
        hidden_2 = self.recurrent_cell_2(lang_input, hidden_2)
        return hidden_1, hidden_2, masks

    def _query_locals(self, data_dict, target_ids, object_masks, include_self=True,
                      overlay_threshold=CONF.TRAIN.OVERLAID_THRESHOLD):
        """target_ids (B) or (B,T) -> 0/1 float mask (B,K) or (B,T,K) of the num_locals nearest proposals."""
        squeeze = target_ids.dim() == 1
        t = target_ids.view(target_ids.shape[0], -1)
        adj, _ = _ext_graph.knn_adjacency(data_dict["bbox_corner"], object_masks, t, self.num_locals,
                                          self.query_mode == "corner", include_self, overlay_threshold)
        return adj.squeeze(1) if squeeze else adj

    def _add_relation_feat(self, data_dict, obj_feats, target_ids):
        """obj_feats (B,K,F) + the target's num_locals relation features scattered onto its neighbours
        (caption_module.py:394-414; edge_feature is indexed with the un-compacted target id, as there).
        target_ids (B) -> (B,K,F);  target_ids (B,T) -> (B,T,K,F)."""
        rel_all = data_dict["edge_feature"]  # (B,K,L,F)
        adjacent_mat = data_dict["adjacent_mat"]  # (B,K,K)
        B, K = adjacent_mat.shape[0], adjacent_mat.shape[1]
        squeeze = target_ids.dim() == 1
        t = target_ids.view(B, -1)
        T = t.shape[1]
        L, Fd = rel_all.shape[2], rel_all.shape[3]
        rel = torch.gather(rel_all, 1, t.view(B, T, 1, 1).expand(B, T, L, Fd))            # (B,T,L,F)
        rows = torch.gather(adjacent_mat, 1, t.view(B, T, 1).expand(B, T, K))              # (B,T,K) 0/1
        # masked_scatter fills the set positions in ascending order with rel[0], rel[1], ...
        rank = (torch.cumsum(rows, 2) - 1).clamp_(0, L - 1).long()                          # (B,T,K)
        scattered = torch.gather(rel, 2, rank.unsqueeze(-1).expand(B, T, K, Fd)) * rows.unsqueeze(-1)
        out = obj_feats.unsqueeze(1) + scattered if obj_feats.dim() == 3 else obj_feats + scattered
        return out.squeeze(1) if squeeze else out

    def forward(self, data_dict, use_tf=True, is_eval=False, max_len=CONF.TRAIN.MAX_DES_LEN):
        if not is_eval:
            return self._forward_sample_batch(data_dict, max_len)
        return self._forward_scene_batch(data_dict, use_tf, max_len)

    def _forward_sample_batch(self, data_dict, max_len=CONF.TRAIN.MAX_DES_LEN, min_iou=CONF.TRAIN.MIN_IOU_THRESHOLD):
        word_embs = data_dict["lang_feat"]     # (B,max_len,emb)
        obj_feats = data_dict["bbox_feature"]  # (B,K,F)
        object_masks = data_dict["bbox_mask"]  # (B,K)
        num_words = _num_words(data_dict)
        B = word_embs.shape[0]
        dev = obj_feats.device

        if self.use_oracle:
            target_ids = data_dict["bbox_idx"]
            target_ious = torch.ones(B, device=dev)
        else:
            target_ids, target_ious = select_target(data_dict)
        target_feats = torch.gather(obj_feats, 1, target_ids.view(B, 1, 1).expand(B, 1, self.feat_size)).squeeze(1)
        valid_masks = object_masks if self.num_locals == -1 else self._query_locals(data_dict, target_ids, object_masks)
        if self.use_relation:
            obj_feats = self._add_relation_feat(data_dict, obj_feats, target_ids)

        # Loop-invariant work hoisted out of the recurrence (same arithmetic up to fp32 summation order):
        #   map_feat(obj_feats)                      -- the reference recomputes it every step (caption_module.py:275)
        #   map_topdown's word / target-feature terms -- W_td [w_t, h2, target] = W_w w_t + W_h h2 + W_f target + b:
        #       W_w w_t for all steps in one GEMM, W_f target + b once; only W_h h2 stays in the loop
        #   classifier(h2)                           -- one (B*T, 512) x (512, V) GEMM after the loop
        T = max(num_words - 1, 1)
        E, H = self.emb_size, self.hidden_size
        # (the Linear layers around the recurrence run on libs2c's fp32 GEMM kernels, lib/linear_simt.py: no library GEMM)
        mapped = linear(obj_feats, self.map_feat.weight)
        w_td, b_td = self.map_topdown[0].weight, self.map_topdown[0].bias
        pre_word = linear(word_embs[:, :T], w_td[:, :E])                   # (B,T,emb)
        pre_tgt = linear(target_feats, w_td[:, E + H:], b_td)              # (B,emb)
        w_td_h = w_td[:, E:E + H]
        # the whole recurrence in one launch (and one for its backward): csrc/caption.cu / caption_grid.cu.  There is
        # no framework-kernel alternative in the product: shapes the kernels do not take raise.
        caption_decoder.require_supported(pre_word, mapped, obj_feats)
        valid_f = (valid_masks != 0).to(torch.float32)
        hid, attn = caption_decoder.topdown_decode(pre_word, pre_tgt, mapped, obj_feats, valid_f, w_td_h,
                                                   self.recurrent_cell_1, self.map_hidd, self.attend,
                                                   self.map_lang[0], self.recurrent_cell_2)
        good_bbox_masks = target_ious > min_iou
        data_dict["lang_cap"] = linear(hid, self.classifier.weight, self.classifier.bias)   # (B,T,V)
        data_dict["pred_ious"] = _masked_mean(target_ious, good_bbox_masks)
        data_dict["topdown_attn"] = attn                        # (B,K,T)
        data_dict["valid_masks"] = valid_masks
        data_dict["good_bbox_masks"] = good_bbox_masks
        return data_dict

    def _forward_scene_batch(self, data_dict, use_tf=False, max_len=CONF.TRAIN.MAX_DES_LEN, chunk=64):
        """Greedy decoding of a caption for EVERY proposal (caption_module.py:502-592): max_len-1 steps.
        Proposals are processed `chunk` at a time (each needs its own (K,F) context once relations are added)."""
        word_embs = data_dict["lang_feat"]
        obj_feats = data_dict["bbox_feature"]
        object_masks = data_dict["bbox_mask"]
        B, K, Fd = obj_feats.shape
        dev = obj_feats.device
        table = _embedding_table(self, dev)
        steps = max_len - 1
        all_out, all_attn, all_valid = [], [], []
        for p0 in range(0, K, chunk):
            P = min(chunk, K - p0)
            tids = torch.arange(p0, p0 + P, device=dev).unsqueeze(0).expand(B, P)
            target_feats = obj_feats[:, p0:p0 + P].reshape(B * P, Fd)
            if self.num_locals == -1:
                valid = object_masks.unsqueeze(1).expand(B, P, K)
            else:
                valid = self._query_locals(data_dict, tids, object_masks)            # (B,P,K)
            all_valid.append(valid)
            if self.use_relation:
                ctx = self._add_relation_feat(data_dict, obj_feats, tids)            # (B,P,K,F)
            else:
                ctx = obj_feats.unsqueeze(1).expand(B, P, K, Fd)
            ctx = ctx.reshape(B * P, K, Fd)
            mapped = self.map_feat(ctx)
            step_masks = valid.reshape(B * P, K, 1)
            h1 = torch.zeros(B * P, self.hidden_size, device=dev)
            h2 = torch.zeros(B * P, self.hidden_size, device=dev)
            step_input = word_embs[:, 0].unsqueeze(1).expand(B, P, self.emb_size).reshape(B * P, self.emb_size)
            outs, attns = [], []
            for _ in range(steps):
                h1, h2, step_mask = self._step(step_input, target_feats, ctx, h1, h2, step_masks, mapped)
                logits = self.classifier(h2)
                outs.append(logits.view(B, P, 1, -1))
                attns.append(step_mask.view(B, P, K, 1))
                step_input = table.index_select(0, logits.argmax(-1))
            all_out.append(torch.cat(outs, dim=2))
            all_attn.append(torch.cat(attns, dim=3))
        data_dict["lang_cap"] = torch.cat(all_out, dim=1)        # (B,K,steps,V)
        data_dict["topdown_attn"] = torch.cat(all_attn, dim=1)   # (B,K,K,steps)
        data_dict["valid_masks"] = torch.cat(all_valid, dim=1)   # (B,K,K)
        return data_dict
