"""The caption-prediction loop of benchmark/predict.py (predict_caption :148-227, decode_caption :132-146) on the
product: EvalStep (the whole inference forward as a CUDA-graph replay), detection loss labels, device-side
parse_predictions (lib/ap_helper.py of this package), then the reference's output structure
{scene_id: [{"caption", "box", "sem_prob", "obj_prob"}]} built from ONE device->host transfer per batch."""
import torch

from ..engine import EvalStep
from .ap_helper import parse_predictions_device

POST_DICT = {"remove_empty_box": True, "use_3d_nms": True, "nms_iou": 0.25, "use_old_type_nms": False, "cls_nms": True,
             "per_class_proposal": True, "conf_thresh": 0.05}   # benchmark/predict.py:161-170


def decode_caption(raw_caption, idx2word):
    """benchmark/predict.py:132-146: "sos w1 w2 ... eos"."""
    decoded = ["sos"]
    for token_idx in raw_caption:
        token = idx2word[str(int(token_idx))]
        decoded.append(token)
        if token == "eos":
            break
    if "eos" not in decoded:
        decoded.append("eos")
    return " ".join(decoded)


class CaptionPredictor(object):
    def __init__(self, model, dataset_config, vocabulary, use_cuda_graph=True):
        self.engine = EvalStep(model, use_cuda_graph=use_cuda_graph)
        self.post = dict(POST_DICT, dataset_config=dataset_config)
        self.idx2word = vocabulary["idx2word"]

    def predict_batch_device(self, data_dict):
        """Everything that runs on the GPU for one batch -> dict of device tensors."""
        out = self.engine.run(data_dict)
        pp = parse_predictions_device(out, self.post)
        final_mask = pp["pred_mask"].long() * out["bbox_mask"].long()          # nms mask x objectness mask (:188-194)
        return {"tokens": out["lang_cap"].argmax(-1), "mask": final_mask, "box": out["bbox_corner"],
                "sem_prob": torch.softmax(out["sem_cls_scores"], -1), "obj_prob": torch.softmax(out["objectness_scores"], -1)}

    def predict_batch(self, data_dict, scene_ids=None):
        dev = self.predict_batch_device(data_dict)
        host = {k: v.cpu() for k, v in dev.items()}
        B, K = host["mask"].shape
        outputs = {}
        for b in range(B):
            scene = []
            for k in range(K):
                if int(host["mask"][b, k]) == 1:
                    scene.append({"caption": decode_caption(host["tokens"][b, k].tolist(), self.idx2word),
                                  "box": host["box"][b, k].numpy().tolist(),
                                  "sem_prob": host["sem_prob"][b, k].numpy().tolist(),
                                  "obj_prob": host["obj_prob"][b, k].numpy().tolist()})
            outputs[scene_ids[b] if scene_ids is not None else b] = scene
        return outputs
