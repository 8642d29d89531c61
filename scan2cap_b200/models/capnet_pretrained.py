"""Mirror of models/capnet_pretrained.py (:9-49): graph + caption on pre-extracted box features
(bbox_feature (B,K,128), bbox_corner (B,K,8,3) f64, bbox_mask (B,K)); K = 128 for mode "gt" (target given by
data_dict["bbox_idx"]) and 256 for "votenet"."""
import torch.nn as nn

from .caption_module import SceneCaptionModule, TopDownSceneCaptionModule
from .graph_module import GraphModule


class CapNet(nn.Module):
    def __init__(self, mode, vocabulary, embeddings, use_topdown=False, num_locals=-1, query_mode="corner",
                 graph_mode="graph_conv", num_graph_steps=0, use_relation=False, graph_aggr="add",
                 use_orientation=False, num_bins=6, use_distance=False, emb_size=300, hidden_size=512):
        super().__init__()
        self.mode = mode
        self.num_graph_steps = num_graph_steps
        num_proposals = self.num_proposals = 128 if mode == "gt" else 256
        use_oracle = mode == "gt"
        if use_relation:
            assert use_topdown
        if num_graph_steps > 0:
            self.graph = GraphModule(128, 128, num_graph_steps, num_proposals, 128, num_locals, query_mode,
                                     graph_mode, return_edge=use_relation, graph_aggr=graph_aggr,
                                     return_orientation=use_orientation, num_bins=num_bins,
                                     return_distance=use_distance)
        if use_topdown:
            self.caption = TopDownSceneCaptionModule(vocabulary, embeddings, emb_size, 128, hidden_size,
                                                     num_proposals, num_locals, query_mode, use_relation, use_oracle)
        else:
            self.caption = SceneCaptionModule(vocabulary, embeddings, emb_size, 128, hidden_size, num_proposals)

    def forward(self, data_dict, use_tf=True, is_eval=False):
        if self.num_graph_steps > 0:
            data_dict = self.graph(data_dict)
        return self.caption(data_dict, use_tf, is_eval)
