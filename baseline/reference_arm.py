"""bench.py --impl reference: the reference's stock code path, timed.

GPU arm: the UNMODIFIED reference Python modules (baseline/_ref: models/capnet.py, lib/pointnet2/*.py,
lib/loss_helper.py ...) over the reference's own CUDA kernels built for sm_100 (oracle/_ref/pointnet2_ref_ext.so),
issued the way lib/solver.py:293-300,376-408 issues a training iteration: every key .to(device) -> model(data_dict)
-> get_scene_cap_loss -> optimizer.zero_grad() -> loss.backward() -> optimizer.step().  CUDA_LAUNCH_BLOCKING is
left unset (scripts/train.py:354 sets it: unset is favourable to the reference).  With N > 1 ranks each rank is an
independent replica (the reference has no multi-GPU mode): no collective.

CPU arm (BASELINE config 1, `capnet_pretrained_cpu`): models/capnet_pretrained.CapNet("votenet", top-down,
relation, orientation, 2 graph steps, 10 locals) -- graph + caption only, bypasses lib/pointnet2 -- on the host
cores with the hard-coded .cuda() calls neutralised, loss of lib/loss_helper_pretrained.py, Adam step.

This module imports neither scan2cap_b200 nor the oracle restatements; the synthetic workload generator
(scan2cap_b200/synthetic.py, numpy only) is loaded BY PATH so that the product package is not imported either.
"""
import importlib.util
import os
import sys
import time

import numpy as np
import torch

from . import shims

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
MODEL_CFG = dict(num_proposal=256, num_locals=10, use_topdown=True, query_mode="center", graph_mode="edge_conv",
                 num_graph_steps=2, use_relation=True, use_orientation=True)


def load_synthetic():
    name = "s2c_synthetic_workload"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "scan2cap_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def assert_clean_process():
    """The reference arm must not have the product or the oracle restatements in its process."""
    bad = [m for m in sys.modules if m == "scan2cap_b200" or m.startswith("scan2cap_b200.") or
           m in ("oracle.ref_model", "oracle.ref_loss", "oracle.native")]
    assert not bad, "reference arm imported %s" % bad
    try:
        with open("/proc/self/maps") as f:
            assert "libs2c.so" not in f.read(), "libs2c.so is mapped in the reference arm's process"
    except OSError:
        pass


class GpuReference(object):
    """Full CapNet training step through the reference's own modules (BASELINE configs 3 / 4)."""

    def __init__(self, input_feature_dim, vocab_size, device, lr=1e-3, weight_decay=1e-5, seed=42):
        ext = shims.load_reference_ext()
        if ext is None:
            raise RuntimeError("oracle/_ref/pointnet2_ref_ext.so is missing (python -m oracle.build_ref)")
        shims.install(ext=ext, cpu=False)
        from data.scannet.model_util_scannet import ScannetDatasetConfig
        from lib.loss_helper import get_scene_cap_loss
        from models.capnet import CapNet
        self.syn = load_synthetic()
        self.DC = ScannetDatasetConfig()
        self.loss_fn = get_scene_cap_loss
        self.device = device
        vocab, emb, _ = self.syn.make_vocabulary(vocab_size)
        self.vocabulary = vocab
        torch.manual_seed(seed)
        self.model = CapNet(self.DC.num_class, vocab, emb, self.DC.num_heading_bin, self.DC.num_size_cluster,
                            self.DC.mean_size_arr, input_feature_dim=input_feature_dim, **MODEL_CFG).to(device)
        self.model.train()
        self.opt = torch.optim.Adam(self.model.parameters(), lr=lr, weight_decay=weight_decay)

    def forward(self, data_dict):
        return self.model(data_dict, use_tf=True)

    def step(self, data_dict):
        """lib/solver.py:376-408.  `data_dict`: tensors on any device."""
        for key in data_dict:
            data_dict[key] = data_dict[key].to(self.device)
        data_dict = self.forward(data_dict)
        data_dict = self.loss_fn(data_dict=data_dict, device=self.device, config=self.DC, weights=None,
                                 detection=True, caption=True, orientation=True, distance=False)
        self.opt.zero_grad()
        data_dict["loss"].backward()
        self.opt.step()
        return data_dict["loss"]


    def predict_batch(self, data_dict, vocabulary):
        """benchmark/predict.py:170-227 for one batch: eval forward with greedy decoding of every proposal, detection
        loss labels, parse_predictions (host-side NMS), decode_caption, the per-scene output lists."""
        from lib.ap_helper import parse_predictions
        post = {"remove_empty_box": True, "use_3d_nms": True, "nms_iou": 0.25, "use_old_type_nms": False, "cls_nms": True,
                "per_class_proposal": True, "conf_thresh": 0.05, "dataset_config": self.DC}
        self.model.eval()
        for key in data_dict:
            data_dict[key] = data_dict[key].to(self.device)
        with torch.no_grad():
            data_dict = self.model(data_dict, False, True)
            data_dict = self.loss_fn(data_dict, self.device, self.DC, weights=None, detection=True, caption=False)
        pred_captions = data_dict["lang_cap"].argmax(-1)
        pred_boxes = data_dict["bbox_corner"]
        parse_predictions(data_dict, post)
        nms_masks = torch.FloatTensor(data_dict["pred_mask"]).type_as(pred_boxes).long() * data_dict["bbox_mask"]
        pred_sem_prob = torch.softmax(data_dict["sem_cls_scores"], dim=-1)
        pred_obj_prob = torch.softmax(data_dict["objectness_scores"], dim=-1)
        idx2word = vocabulary["idx2word"]
        outputs = {}
        for batch_id in range(pred_captions.shape[0]):
            scene_outputs = []
            for object_id in range(pred_captions.shape[1]):
                if nms_masks[batch_id, object_id] == 1:
                    decoded = ["sos"]
                    for token_idx in pred_captions[batch_id, object_id]:
                        token = idx2word[str(token_idx.item())]
                        decoded.append(token)
                        if token == "eos":
                            break
                    if "eos" not in decoded:
                        decoded.append("eos")
                    scene_outputs.append({"caption": " ".join(decoded),
                                          "box": pred_boxes[batch_id, object_id].cpu().detach().numpy().tolist(),
                                          "sem_prob": pred_sem_prob[batch_id, object_id].cpu().detach().numpy().tolist(),
                                          "obj_prob": pred_obj_prob[batch_id, object_id].cpu().detach().numpy().tolist()})
            outputs[batch_id] = scene_outputs
        self.model.train()
        return outputs


class CpuPretrainedReference(object):
    """BASELINE config 1: the reference's CPU-runnable case."""

    def __init__(self, vocab_size=3500, seed=42):
        shims.install(ext=None, cpu=True)
        from lib.loss_helper_pretrained import get_loss
        from models.capnet_pretrained import CapNet
        self.syn = load_synthetic()
        vocab, emb, _ = self.syn.make_vocabulary(vocab_size)
        torch.manual_seed(seed)
        self.model = CapNet("votenet", vocab, emb, use_topdown=True, num_locals=10, query_mode="center",
                            graph_mode="edge_conv", num_graph_steps=2, use_relation=True, use_orientation=True)
        self.model.train()
        self.loss_fn = get_loss
        self.opt = torch.optim.Adam(self.model.parameters(), lr=1e-3, weight_decay=1e-5)
        self.vocab_size = vocab_size

    def step(self, data_dict):
        data_dict = self.model(data_dict, use_tf=True)
        data_dict = self.loss_fn(data_dict=data_dict, mode="votenet", orientation=True)
        self.opt.zero_grad()
        data_dict["loss"].backward()
        self.opt.step()
        return data_dict["loss"]


def time_cpu_pretrained(budget_s=12.0, max_steps=200, vocab_size=3500, threads=None):
    """scenes/s of config 1 (B = 1 scene per step) on the host cores; a bounded sample (>= `budget_s` seconds)."""
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    ref = CpuPretrainedReference(vocab_size)
    d = ref.syn.make_pretrained_data_dict(1, num_proposals=256, num_valid=64, num_vocabs=vocab_size, seed=7)
    host = {k: torch.from_numpy(v) for k, v in d.items()}
    ref.step({k: v.clone() for k, v in host.items()})  # warm-up (allocator, thread pool)
    steps, t0 = 0, time.perf_counter()
    while True:
        ref.step({k: v.clone() for k, v in host.items()})
        steps += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or steps >= max_steps:
            break
    return {"value": steps / dt, "unit": "scenes/s", "cores": os.cpu_count(), "kind": "reference",
            "threads": {"torch": torch.get_num_threads()}, "config": "BASELINE configs[0]: capnet_pretrained "
            "(votenet mode, graph+caption only, bypasses lib/pointnet2), batch 1, 256 proposals (64 valid), "
            "20-token caption, V=%d, CPU" % vocab_size,
            "implementation": "the reference's own models/capnet_pretrained.py + lib/loss_helper_pretrained.py "
            "(baseline/_ref), .cuda() neutralised, PyG propagate restated (baseline/shims.py)",
            "sample": "%d training step(s) of 1 scene, %.1f s of host time" % (steps, dt)}


if __name__ == "__main__":
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu-pretrained", action="store_true")
    ap.add_argument("--budget", type=float, default=12.0)
    a = ap.parse_args()
    if a.cpu_pretrained:
        out = time_cpu_pretrained(budget_s=a.budget)
        assert_clean_process()
        print(json.dumps(out), flush=True)
