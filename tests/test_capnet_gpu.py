"""GPU parity of the product CapNet (scan2cap_b200, fused / batched / sync-free, libs2c kernels through the
C ABI) against the oracle restatement of the reference (oracle/ref_model.py + ref_loss.py driving the
reference's own CUDA kernels when oracle/_ref is present, else the C oracle), same weights, same inputs -- at the
shapes the benchmark runs (BASELINE configs[2] "c3": B=8, N=40 000, 7 floats; configs[3] "c4": B=4, N=40 000, 135
floats with the shipped XYZ_MULTIVIEW_NORMAL VoteNet checkpoint mounted as scripts/train.py:84-104 does) and at two
small shapes.

Bars (north_star): integer outputs (FPS indices, neighbour lists via the masks/adjacency, kNN edges) bit-exact;
float features / logits / losses within 1e-3 relative (of the tensor's max magnitude); BatchNorm running statistics
within 1e-3; gradients: relative L2 <= 1e-3 PER PARAMETER with the ReLU / max-pool decisions of the two sides
compared and the groups whose decisions flipped excluded from both backward passes (tests/parity_utils.py; the
flip count is printed).

FPS on the VOTE coordinates is discontinuous in its float input: a 1e-6 difference in vote_xyz between two fp32
evaluation orders can flip a pick and with it the whole proposal set.  When that happens the run is repeated with the
product's vote-aggregation sampling forced to the oracle's picks (after checking that our FPS kernel reproduces those
picks bit-exactly from the oracle's votes), and EVERYTHING is still compared end to end -- outputs, loss, every
gradient.  Which branch each case took is recorded in BRANCHES (printed, and written to gpurun_out/) and
test_parity_branches fails if more than two cases needed the forced branch."""
import copy
import json
import os
import warnings

import numpy as np
import pytest
import torch

import parity_utils as PU
from scan2cap_b200 import synthetic
from scan2cap_b200.data.scannet.model_util_scannet import ScannetDatasetConfig

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-3
BRANCHES = {}

INT_KEYS = ["sa1_inds", "sa2_inds", "fp2_inds", "aggregated_vote_inds", "bbox_mask", "bbox_sems", "num_edge_source",
            "num_edge_target", "good_bbox_masks", "object_assignment", "objectness_label"]
EXACT_FLOAT_KEYS = ["sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz", "adjacent_mat", "edge_index", "valid_masks"]
FLOAT_KEYS = ["sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features", "vote_xyz",
              "vote_features", "aggregated_vote_xyz", "aggregated_vote_features", "objectness_scores", "center",
              "heading_scores", "size_scores", "size_residuals", "sem_cls_scores", "bbox_corner", "bbox_feature",
              "edge_feature", "edge_orientations", "edge_distances", "lang_cap", "topdown_attn", "pred_ious", "loss",
              "vote_loss", "objectness_loss", "box_loss", "sem_cls_loss", "cap_loss", "ori_loss", "dist_loss", "cap_acc",
              "ori_acc", "obj_acc"]
PRE_INT = ["sa1_inds", "sa2_inds", "fp2_inds"]
PRE_EXACT = ["sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz"]
PRE_FLOAT = ["sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features", "vote_xyz", "vote_features"]

_rel = PU.rel


def _models(query_mode, C, V, seed=0, checkpoint=None):
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
    if checkpoint is not None:
        # scripts/train.py:84-104: the pre-trained VoteNet (backbone, voting, proposal) is loaded with strict=False
        # into a no_caption CapNet and mounted; here: its tensors are loaded straight into the full model
        sd = torch.load(checkpoint, map_location=DEV)
        missing, unexpected = ours.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.startswith(("graph.", "caption.")) for k in missing), (missing, unexpected)
    ref.load_state_dict(ours.state_dict(), strict=True)  # identical key set = the checkpoint contract
    return ours, ref, DC


def _data(B, N, V, seed, use_normal=True, use_multiview=False):
    d = synthetic.make_data_dict(B, N, use_normal=use_normal, use_multiview=use_multiview, num_vocabs=V, seed=seed)
    return {k: torch.from_numpy(v).to(DEV) for k, v in d.items()}


def _clone(d):
    return {k: v.clone() for k, v in d.items()}


def _check_outputs(o, r, int_keys, exact_keys, float_keys):
    for k in int_keys:
        assert torch.equal(o[k].long(), r[k].long()), "integer output %s differs" % k
    for k in exact_keys:
        a, b = o[k].double(), r[k].double()
        if k == "adjacent_mat":
            # rows of invalid proposals are computed but never used (graph_module.py:266 keeps the valid rows / columns
            # only); they depend on near-ties of float64 centre distances whose inputs differ by 1e-7 between the sides
            v = (r["bbox_mask"] > 0).double().unsqueeze(-1)
            a, b = a * v, b * v
        if not torch.equal(a, b):
            diff = (a != b)
            info = "%d of %d entries" % (int(diff.sum()), diff.numel())
            if k == "adjacent_mat":
                rows = diff.any(-1)
                info += "; rows differing per scene %s; valid proposals per scene %s; differing rows that are valid %s" % (
                    rows.sum(1).tolist(), r["bbox_mask"].sum(1).tolist(), (rows & (r["bbox_mask"] > 0)).sum(1).tolist())
            raise AssertionError("index-like output %s differs: %s" % (k, info))
    worst = {}
    for k in float_keys:
        assert o[k].shape == r[k].shape, k
        worst[k] = _rel(o[k], r[k])
    bad = {k: v for k, v in worst.items() if not v < RTOL}
    assert not bad, "float outputs beyond %g: %s" % (RTOL, bad)
    return max(worst.values())


CASES = {
    # name: (query_mode, B, N, use_multiview, vocabulary, checkpoint, data seed)
    "small_center": ("center", 2, 8000, False, 150, None, 11),
    # (The two point-wise heads keep a residual the exclusion cannot remove: BatchNorm1d couples the rows of a batch, so
    #  a flipped ReLU in an excluded row still reaches every upstream gradient through the batch statistics, by about
    #  1 / (B * 256) per flip for the proposal head -- measured with tools/parity_matrix.py: B = 1 cases exceed the bar
    #  on most seeds, B = 2 on some, B >= 4 on none.  The small cases are therefore B = 2 / N = 8000; B = 1 is covered
    #  forward-only (test_configs_gpu.py) and by the BatchNorm-free config 1.)
    "small_corner": ("corner", 2, 8000, False, 150, None, 14),
    "c3_B8_N40k_C4": ("center", 8, 40000, False, 3500, None, 42),
    "c4_B4_N40k_C132_ckpt": ("center", 4, 40000, True, 3500, "PRETRAIN_VOTENET_XYZ_MULTIVIEW_NORMAL", 42),
}


@pytest.mark.parametrize("case", list(CASES))
def test_capnet_forward_backward_parity(case):
    from oracle import ref_loss as RL
    from scan2cap_b200.lib.loss_helper import get_scene_cap_loss
    from scan2cap_b200.lib.pointnet2 import _ext
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    query_mode, B, N, use_mv, V, ckpt, seed = CASES[case]
    ckpt_path = None
    if ckpt is not None:
        ckpt_path = PU.checkpoint_path(ckpt)
        if ckpt_path is None:
            pytest.skip("shipped checkpoint %s not installed under baseline/_ref/pretrained" % ckpt)
    C = 4 + (128 if use_mv else 0)
    ours, ref, DC = _models(query_mode, C, V, checkpoint=ckpt_path)
    data = _data(B, N, V, seed=seed, use_multiview=use_mv)
    state = copy.deepcopy(ours.state_dict())
    with torch.no_grad():
        probe = ours(_clone(data))
    data["ref_box_corner_label"] = probe["bbox_corner"][:, 7].clone()
    data["ref_box_corner_label"][-1] += 50.0  # last scene: no good box -> its caption loss is masked out
    del probe
    ours.train(); ref.train()
    flags = (True, True, True, True)
    with PU.DecisionTracker(ours, ref) as tracker:
        # The oracle runs with cuDNN disabled (ATen's native conv / batch-norm kernels): cuDNN's fp32 weight-gradient
        # kernels for 1x1 convolutions over 10^5..10^6 positions deviate 3e-3..2e-2 from a float64 evaluation on
        # B200 (profiles/r01_gradient_accuracy.txt) whereas the GEMM formulation is at 1e-6, so cuDNN's gradients
        # cannot serve as the yardstick.  bench.py --impl reference keeps cuDNN on (stock path).
        with torch.backends.cudnn.flags(enabled=False):
            r = RL.get_scene_cap_loss(ref(_clone(data)), DEV, DC, None, *flags)

        def run_ours(forced_inds=None):
            ours.load_state_dict(state)   # undo BatchNorm running-statistics updates of earlier forward passes
            ours.zero_grad()
            tracker.fused_mlp.CAPTURE = []
            tracker.graph_module.CAPTURE = []
            tracker.caption_decoder.CAPTURE = []
            va = ours.proposal.vote_aggregation
            if forced_inds is not None:
                orig = va.forward
                va.forward = lambda xyz, features=None, inds=None, sampled_xyz=None: orig(xyz, features, inds=forced_inds)
            try:
                return get_scene_cap_loss(ours(_clone(data)), DEV, DC, None, *flags)
            finally:
                if forced_inds is not None:
                    del va.forward

        o = run_ours()
        # ---- everything up to the votes: FPS / ball-query indices bit-exact, features within tolerance
        _check_outputs(o, r, PRE_INT, PRE_EXACT, PRE_FLOAT)
        branch = "end-to-end"
        if not torch.equal(o["aggregated_vote_inds"].long(), r["aggregated_vote_inds"].long()):
            # our FPS kernel on the ORACLE's votes reproduces the oracle's picks: the flip came from input rounding
            mine = _ext.furthest_point_sampling(r["vote_xyz"].detach().contiguous(), 256)
            assert torch.equal(mine.long(), r["aggregated_vote_inds"].long()), "vote FPS differs on identical votes"
            nflip = int((o["aggregated_vote_inds"].long() != r["aggregated_vote_inds"].long()).any(1).sum())
            branch = "vote-FPS picks forced to the oracle's (%d of %d scenes flipped)" % (nflip, B)
            o = run_ours(forced_inds=r["aggregated_vote_inds"].int().contiguous())
            _check_outputs(o, r, PRE_INT, PRE_EXACT, PRE_FLOAT)
        worst_out = _check_outputs(o, r, INT_KEYS, EXACT_FLOAT_KEYS, FLOAT_KEYS)
        flipped = tracker.resolve(o, r)
        tracker.exclude_head_rows(o, r)
        tracker.resolve_graph_and_caption(o, r)
        loss_key = os.environ.get("S2C_PARITY_LOSS", "loss")   # diagnostic: back-propagate a single loss term
        o[loss_key].backward()
        with torch.backends.cudnn.flags(enabled=False):
            r[loss_key].backward()
    worst_grad = PU.check_grads_per_parameter(ours, ref, label="[%s] " % case)
    rb = dict(ref.named_buffers())
    for n1, b1 in ours.named_buffers():
        if "running" in n1:
            assert _rel(b1, rb[n1]) < RTOL, n1
    BRANCHES[case] = {"branch": branch, "flipped_groups": tracker.report(), "worst_output_rel": worst_out,
                      "worst_param_grad_rel_l2": worst_grad}
    msg = "parity[%s]: %s; decision flips excluded: %s (total %d); worst output %.2e, worst parameter gradient %.2e" % (
        case, branch, tracker.report(), flipped, worst_out, worst_grad)
    print(msg)
    warnings.warn(msg)   # shows up in the pytest summary of a green run (the driver's log)


def test_parity_branches():
    """Record which branch every parity case took; more than two forced-pick cases would mean the end-to-end
    comparison is the exception rather than the rule."""
    if not BRANCHES:
        pytest.skip("no parity case ran in this session")
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_branches.json"), "w") as f:
            json.dump(BRANCHES, f, indent=1)
    except OSError:
        pass
    forced = [c for c, v in BRANCHES.items() if v["branch"] != "end-to-end"]
    assert len(forced) <= 2, "vote-FPS picks had to be forced in %s" % forced


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


def test_graph_conv_mode_matches_restatement():
    """graph_mode="graph_conv" (GCNConv, graph_module.py:136): batched masked product implementation vs the oracle's
    per-scene compacted restatement (PyG internals themselves: parity unpinned)."""
    from oracle import ref_model as R
    from scan2cap_b200.models.graph_module import GraphModule
    torch.manual_seed(0)
    args = (128, 128, 2, 256, 128, 10, "center", "graph_conv", False, "add", False, 6, False)
    ours, ref = GraphModule(*args).to(DEV), R.GraphModule(*args).to(DEV)
    ref.load_state_dict(ours.state_dict(), strict=True)
    B, K = 2, 256
    g = torch.Generator().manual_seed(1)
    centers = torch.rand(B, K, 3, generator=g, dtype=torch.float64) * 6
    sizes = torch.rand(B, K, 3, generator=g, dtype=torch.float64) * 0.5 + 0.1
    corners = torch.from_numpy(synthetic.box_corners(centers.numpy(), sizes.numpy())).to(DEV)
    mask = (torch.rand(B, K, generator=g) > 0.4).long()
    feats = torch.randn(B, K, 128, generator=g)
    fo = feats.clone().to(DEV).requires_grad_(True)
    fr = feats.clone().to(DEV).requires_grad_(True)
    o = ours({"bbox_feature": fo, "bbox_mask": mask.to(DEV), "bbox_corner": corners})
    r = ref({"bbox_feature": fr, "bbox_mask": mask.to(DEV), "bbox_corner": corners})
    assert torch.equal(o["adjacent_mat"], r["adjacent_mat"])
    assert _rel(o["bbox_feature"], r["bbox_feature"]) < 1e-5
    gout = torch.randn(B, K, 128, generator=g).to(DEV)
    (o["bbox_feature"] * gout).sum().backward()
    (r["bbox_feature"] * gout).sum().backward()
    assert _rel(fo.grad, fr.grad) < 1e-5
    for (n, po), (_, pr) in zip(ours.named_parameters(), ref.named_parameters()):
        assert _rel(po.grad, pr.grad) < 1e-5, n
