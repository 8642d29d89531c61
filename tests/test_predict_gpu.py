"""Inference path: device-side post-processing vs the reference's OWN lib/ap_helper.py / utils/nms.py
(the verbatim copy under baseline/_ref, skipped where absent), EvalStep graph replay vs eager, predictor output."""
import os

import numpy as np
import pytest
import torch

from scan2cap_b200 import synthetic
from scan2cap_b200.data.scannet.model_util_scannet import ScannetDatasetConfig

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "lib", "ap_helper.py"))


def _reference_modules():
    from baseline import shims
    shims.install(ext=None, cpu=False)
    import lib.ap_helper as ref_ap
    import utils.nms as ref_nms
    from data.scannet.model_util_scannet import ScannetDatasetConfig as RefDC, extract_pc_in_box3d
    return ref_ap, ref_nms, RefDC, extract_pc_in_box3d


def _random_boxes(rng, B, K):
    c = rng.uniform(-3, 3, (B, K, 3))
    s = rng.uniform(0.3, 2.0, (B, K, 3))
    return np.concatenate([c - s / 2, c + s / 2], -1)


@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref not installed")
@pytest.mark.parametrize("same_cls,old_type", [(True, False), (False, False), (True, True)])
def test_nms3d_matches_reference_numpy(same_cls, old_type):
    from scan2cap_b200.lib.ap_helper import nms3d
    _, ref_nms, _, _ = _reference_modules()
    rng = np.random.default_rng(0)
    B, K = 3, 256
    boxes = _random_boxes(rng, B, K)
    score = rng.random((B, K))
    cls = rng.integers(0, 4, (B, K))
    valid = (rng.random((B, K)) > 0.2).astype(np.int32)
    valid[2] = 0
    valid[2, 5] = 1   # a scene with a single valid box
    keep = nms3d(torch.from_numpy(boxes).to(DEV), torch.from_numpy(score).to(DEV), torch.from_numpy(cls).to(DEV),
                 torch.from_numpy(valid).to(DEV), 0.25, old_type, same_cls).cpu().numpy()
    for b in range(B):
        ids = np.where(valid[b] == 1)[0]
        arr = np.concatenate([boxes[b, ids], score[b, ids, None], cls[b, ids, None].astype(np.float64)], 1)
        pick = (ref_nms.nms_3d_faster_samecls(arr, 0.25, old_type) if same_cls else
                ref_nms.nms_3d_faster(arr[:, :7], 0.25, old_type))
        want = np.zeros(K, np.int32)
        want[ids[pick]] = 1
        np.testing.assert_array_equal(keep[b], want)


@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref not installed")
def test_points_in_boxes_and_parse_predictions_match_reference():
    from scan2cap_b200.lib.ap_helper import parse_predictions, points_in_boxes_count
    ref_ap, _, RefDC, extract_pc_in_box3d = _reference_modules()
    DC = RefDC()
    rng = np.random.default_rng(1)
    B, K, N = 2, 256, 6000
    pc, _ = synthetic.make_point_clouds(B, N, use_normal=True, use_height=True, seed=5)
    pcs = torch.from_numpy(pc).to(DEV)
    # decoded boxes around real geometry so that some are empty and some are not
    center = torch.from_numpy(pc[:, rng.permutation(N)[:K], :3] + rng.normal(0, 0.2, (B, K, 3)).astype(np.float32)).to(DEV)
    NS, NC = DC.num_size_cluster, DC.num_class
    end_points = {
        "point_clouds": pcs, "center": center,
        "heading_scores": torch.zeros(B, K, 1, device=DEV), "heading_residuals": torch.zeros(B, K, 1, device=DEV),
        "size_scores": torch.from_numpy(rng.normal(0, 1, (B, K, NS)).astype(np.float32)).to(DEV),
        "size_residuals": torch.from_numpy(rng.normal(0, 0.2, (B, K, NS, 3)).astype(np.float32)).to(DEV),
        "sem_cls_scores": torch.from_numpy(rng.normal(0, 1, (B, K, NC)).astype(np.float32)).to(DEV),
        "objectness_scores": torch.from_numpy(rng.normal(0, 2, (B, K, 2)).astype(np.float32)).to(DEV),
    }
    cfg = {"remove_empty_box": True, "use_3d_nms": True, "nms_iou": 0.25, "use_old_type_nms": False, "cls_nms": True,
           "per_class_proposal": True, "conf_thresh": 0.05, "dataset_config": DC}
    ours_ep, ref_ep = dict(end_points), dict(end_points)
    mine = parse_predictions(ours_ep, cfg)
    want = ref_ap.parse_predictions(ref_ep, cfg)
    np.testing.assert_array_equal(ours_ep["pred_mask"], ref_ep["pred_mask"])
    assert 0 < ours_ep["pred_mask"].sum() < B * K
    assert [len(a) for a in mine] == [len(b) for b in want]
    for a, b in zip(mine, want):
        for (ca, ba, sa), (cb, bb, sb) in zip(a, b):
            assert ca == cb and abs(sa - sb) <= 1e-6 * max(abs(sb), 1e-3)
            np.testing.assert_allclose(ba, bb, rtol=0, atol=1e-6)
    # the point count itself, against the reference's Delaunay-hull test, box by box
    from scan2cap_b200.lib.ap_helper import parse_predictions_device
    dev = parse_predictions_device(dict(end_points), cfg)
    corners = dev["corners"].cpu().numpy()
    boxes = torch.cat([dev["corners"].amin(2), dev["corners"].amax(2)], -1)
    cnt = points_in_boxes_count(pcs, boxes).cpu().numpy()
    for j in range(0, K, 9):
        inside, _ = extract_pc_in_box3d(pc[0, :, :3], corners[0, j])
        assert len(inside) == cnt[0, j], j


def test_evalstep_graph_matches_eager_and_predictor_output():
    from scan2cap_b200.engine import EvalStep
    from scan2cap_b200.lib.predict import CaptionPredictor
    from scan2cap_b200.models.capnet import CapNet
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    DC = ScannetDatasetConfig()
    V = 80
    vocab, emb, _ = synthetic.make_vocabulary(V)
    torch.manual_seed(0)
    model = CapNet(DC.num_class, vocab, emb, DC.num_heading_bin, DC.num_size_cluster, DC.mean_size_arr,
                   input_feature_dim=4, num_proposal=256, num_locals=10, use_topdown=True, query_mode="corner",
                   graph_mode="edge_conv", num_graph_steps=2, use_relation=True, use_orientation=True).to(DEV).eval()
    d = synthetic.make_data_dict(2, 8000, use_normal=True, num_vocabs=V, seed=3)
    data = {k: torch.from_numpy(v).to(DEV) for k, v in d.items()}
    eager = EvalStep(model, use_cuda_graph=False).run(dict(data))
    graph_engine = EvalStep(model, use_cuda_graph=True)
    for _ in range(2):   # capture, then a pure replay
        out = graph_engine.run(dict(data))
    assert out["lang_cap"].shape == (2, 256, 29, V)
    assert torch.equal(out["lang_cap"].argmax(-1), eager["lang_cap"].argmax(-1))
    assert torch.equal(out["bbox_mask"], eager["bbox_mask"])
    assert float((out["lang_cap"] - eager["lang_cap"]).abs().max()) <= 1e-4 * float(eager["lang_cap"].abs().max())
    pred = CaptionPredictor(model, DC, vocab).predict_batch(dict(data), scene_ids=["scene_a", "scene_b"])
    assert set(pred) == {"scene_a", "scene_b"}
    for scene in pred.values():
        for obj in scene:
            assert obj["caption"].startswith("sos") and obj["caption"].endswith("eos")
            assert np.asarray(obj["box"]).shape == (8, 3) and len(obj["sem_prob"]) == DC.num_class and len(obj["obj_prob"]) == 2
