"""GPU parity at the remaining BASELINE.json configurations:

  configs[1]  VoteNet backbone forward only, 1 x 40 000 points, SHIPPED checkpoints (PRETRAIN_VOTENET_XYZ: C=1,
              PRETRAIN_VOTENET_XYZ_MULTIVIEW_NORMAL: C=132) and the pure-xyz variant: FPS indices of all levels, the
              ball-query neighbour lists of all five grouping stages and both three_nn index tensors BIT-EXACT against
              the reference's own kernels (oracle/_ref) / the C oracle; features within 1e-3 with trained weights.
  configs[0]  capnet_pretrained("votenet") -- graph + caption on pre-extracted box features, batch 1, 64 valid of 256
              proposals -- product vs oracle: kNN edges bit-exact, logits / losses 1e-3, per-parameter gradients 1e-3.
"""
import numpy as np
import pytest
import torch

import parity_utils as PU
from scan2cap_b200 import synthetic
from scan2cap_b200.data.scannet.model_util_scannet import ScannetDatasetConfig

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-3


def _detector_pair(C, checkpoint):
    from conftest import load_reference_ext
    from oracle import ref_model as R
    from scan2cap_b200.models.capnet import CapNet
    R.set_backend(load_reference_ext())
    DC = ScannetDatasetConfig()
    vocab, emb, _ = synthetic.make_vocabulary(20)
    torch.manual_seed(1)
    args = (DC.num_class, vocab, emb, DC.num_heading_bin, DC.num_size_cluster, DC.mean_size_arr)
    ours = CapNet(*args, input_feature_dim=C, num_proposal=256, no_caption=True).to(DEV)
    ref = R.CapNet(*args, input_feature_dim=C, num_proposal=256, no_caption=True).to(DEV)
    if checkpoint is not None:
        ours.load_state_dict(torch.load(checkpoint, map_location=DEV), strict=True)   # "All keys matched"
    ref.load_state_dict(ours.state_dict(), strict=True)
    return ours.eval(), ref.eval()


@pytest.mark.parametrize("ckpt,use_normal,use_mv,use_height", [
    ("PRETRAIN_VOTENET_XYZ", False, False, True),                 # configs[1] as written: xyz + height, C = 1
    ("PRETRAIN_VOTENET_XYZ_MULTIVIEW_NORMAL", True, True, True),  # C = 132: unaligned 540-byte rows of point_clouds
    (None, False, False, False),                                  # --no_height: pure (1, 40000, 3), random init
])
def test_backbone_forward_40k_indices_bit_exact(ckpt, use_normal, use_mv, use_height, ref_ext, oracle, ext):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    path = None
    if ckpt is not None:
        path = PU.checkpoint_path(ckpt)
        if path is None:
            pytest.skip("shipped checkpoint %s not installed under baseline/_ref/pretrained" % ckpt)
    pc, _ = synthetic.make_point_clouds(1, 40000, use_normal=use_normal, use_multiview=use_mv, use_height=use_height,
                                        seed=42)
    C = pc.shape[-1] - 3
    ours, ref = _detector_pair(C, path)
    data = {"point_clouds": torch.from_numpy(pc).to(DEV)}
    with torch.no_grad():
        o = ours(dict(data))
        r = ref(dict(data))
    for k in ("sa1_inds", "sa2_inds", "fp2_inds"):
        assert torch.equal(o[k].long(), r[k].long()), k
    for k in ("sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz", "fp2_xyz"):
        assert torch.equal(o[k], r[k]), k
    for k in ("sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features", "vote_xyz", "vote_features"):
        assert PU.rel(o[k], r[k]) < RTOL, (k, PU.rel(o[k], r[k]))
    # Proposal stage.  FPS on the VOTE coordinates is discontinuous in its float input (a 1e-6 difference in vote_xyz can
    # swap two picks), so it is checked from IDENTICAL votes: the product's proposal module on the oracle's votes must
    # reproduce the oracle's picks bit-exactly, and everything computed from them within tolerance.
    if not torch.equal(o["aggregated_vote_inds"].long(), r["aggregated_vote_inds"].long()):
        print("vote-FPS picks differ end to end (float input rounding); compared from the oracle's votes")
    with torch.no_grad():
        o = ours.proposal(r["vote_xyz"].contiguous(), r["vote_features"].contiguous(), {})
    assert torch.equal(o["aggregated_vote_inds"].long(), r["aggregated_vote_inds"].long()), "vote FPS on identical votes"
    for k in ("aggregated_vote_xyz", "aggregated_vote_features", "objectness_scores", "center", "size_scores",
              "sem_cls_scores", "bbox_corner"):
        assert PU.rel(o[k], r[k]) < RTOL, (k, PU.rel(o[k], r[k]))
    assert torch.equal(o["bbox_mask"], r["bbox_mask"])

    # the neighbour lists of every grouping stage and both three_nn index tensors, product kernels vs the
    # reference's kernels (or the C oracle) on the SAME coordinates (sa*_xyz are bit-identical, checked above)
    xyz0 = data["point_clouds"][..., :3].contiguous()
    stages = [(xyz0, r["sa1_xyz"], 0.2, 64), (r["sa1_xyz"], r["sa2_xyz"], 0.4, 32), (r["sa2_xyz"], r["sa3_xyz"], 0.8, 16),
              (r["sa3_xyz"], r["sa4_xyz"], 1.2, 16), (r["vote_xyz"].contiguous(), r["aggregated_vote_xyz"].contiguous(), 0.3, 16)]
    for i, (xyz, new_xyz, radius, ns) in enumerate(stages):
        mine = ext.ball_query(new_xyz, xyz, radius, ns)
        if ref_ext is not None:
            want = ref_ext.ball_query(new_xyz, xyz, radius, ns)
        else:
            want = torch.from_numpy(oracle.ball_query(new_xyz.cpu().numpy(), xyz.cpu().numpy(), radius, ns)).to(DEV)
        assert torch.equal(mine, want), "ball_query idx of grouping stage %d differs" % i
    for unknown, known in ((r["sa3_xyz"], r["sa4_xyz"]), (r["sa2_xyz"], r["sa3_xyz"])):
        d_m, i_m = ext.three_nn(unknown.contiguous(), known.contiguous())
        if ref_ext is not None:
            d_w, i_w = ref_ext.three_nn(unknown.contiguous(), known.contiguous())
        else:
            d_w, i_w = (torch.from_numpy(a).to(DEV) for a in oracle.three_nn(unknown.cpu().numpy(), known.cpu().numpy()))
        assert torch.equal(i_m, i_w) and torch.equal(d_m, d_w)
    # FPS of every level straight through the _ext surface (40000 -> 2048 -> 1024 -> 512 -> 256)
    cur = xyz0
    for m in (2048, 1024, 512, 256):
        mine = ext.furthest_point_sampling(cur, m)
        want = (ref_ext.furthest_point_sampling(cur, m) if ref_ext is not None else
                torch.from_numpy(oracle.furthest_point_sampling(cur.cpu().numpy(), m)).to(DEV))
        assert torch.equal(mine, want), "FPS %d -> %d" % (cur.shape[1], m)
        cur = torch.gather(cur, 1, want.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()


@pytest.mark.parametrize("query_mode,num_valid", [("center", 64), ("corner", 40)])
def test_capnet_pretrained_config1_parity(query_mode, num_valid):
    """BASELINE configs[0] on the GPU: models/capnet_pretrained.py:35-49 (graph + caption only)."""
    from oracle import ref_model as R
    from scan2cap_b200.lib.loss_helper_pretrained import get_loss
    from scan2cap_b200.models.capnet_pretrained import CapNet
    torch.backends.cuda.matmul.allow_tf32 = False
    V = 3500
    vocab, emb, _ = synthetic.make_vocabulary(V)
    cfg = dict(use_topdown=True, num_locals=10, query_mode=query_mode, graph_mode="edge_conv", num_graph_steps=2,
               use_relation=True, use_orientation=True)
    torch.manual_seed(5)
    ours = CapNet("votenet", vocab, emb, **cfg).to(DEV)
    ref = R.CapNetPretrained("votenet", vocab, emb, **cfg).to(DEV)
    ref.load_state_dict(ours.state_dict(), strict=True)
    d = synthetic.make_pretrained_data_dict(1, num_proposals=256, num_valid=num_valid, num_vocabs=V, seed=9, lang_len=20)
    data = {k: torch.from_numpy(v).to(DEV) for k, v in d.items()}
    ours.train(); ref.train()
    o = get_loss(ours({k: v.clone() for k, v in data.items()}), mode="votenet", orientation=True)
    r = get_loss(ref({k: v.clone() for k, v in data.items()}), mode="votenet", orientation=True)
    assert o["lang_cap"].shape == r["lang_cap"].shape == (1, 19, V)
    for k in ("num_edge_source", "num_edge_target", "good_bbox_masks"):
        assert torch.equal(o[k].long(), r[k].long()), k
    assert int(o["num_edge_source"][0]) == num_valid and int(o["num_edge_target"][0]) == 10
    for k in ("adjacent_mat", "edge_index", "valid_masks"):
        assert torch.equal(o[k].double(), r[k].double()), k
    for k in ("bbox_feature", "edge_feature", "edge_orientations", "lang_cap", "topdown_attn", "pred_ious", "loss",
              "cap_loss", "ori_loss", "cap_acc", "ori_acc"):
        assert PU.rel(o[k], r[k]) < RTOL, (k, PU.rel(o[k], r[k]))
    o["loss"].backward()
    r["loss"].backward()
    PU.check_grads_per_parameter(ours, ref, label="[config1 %s] " % query_mode)


def test_mask_votenet_and_encoder_parity():
    """MaskVoteNet (vote aggregation over a 5 m ball, nsample = 512, one proposal per scene;
    models/mask_votenet.py:145-153) and PointnetEncoder (models/encoder_module.py) on the same kernels: forward
    outputs and every parameter gradient against the oracle restatement."""
    from conftest import load_reference_ext
    from oracle import ref_model as R
    from scan2cap_b200.models.encoder_module import PointnetEncoder
    from scan2cap_b200.models.mask_votenet import MaskVoteNet
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    R.set_backend(load_reference_ext())
    DC = ScannetDatasetConfig()
    pc, _ = synthetic.make_point_clouds(4, 8000, use_normal=True, use_height=True, seed=3)
    data = {"point_clouds": torch.from_numpy(pc).to(DEV)}
    torch.manual_seed(2)
    ours = MaskVoteNet(DC.num_class, DC.num_heading_bin, DC.num_size_cluster, DC.mean_size_arr, input_feature_dim=4).to(DEV)
    ref = R.MaskVoteNet(DC.num_class, DC.num_heading_bin, DC.num_size_cluster, DC.mean_size_arr, input_feature_dim=4).to(DEV)
    ref.load_state_dict(ours.state_dict(), strict=True)
    ours.train(); ref.train()
    o = ours(dict(data))
    with torch.backends.cudnn.flags(enabled=False):
        r = ref(dict(data))
    assert o["center"].shape == (4, 1, 3) and o["sem_cls_scores"].shape == (4, 1, DC.num_class)
    assert torch.equal(o["aggregated_vote_inds"].long(), r["aggregated_vote_inds"].long())
    for k in ("vote_xyz", "vote_features", "aggregated_vote_features", "center", "size_scores",
              "size_residuals", "sem_cls_scores"):
        assert PU.rel(o[k], r[k]) < RTOL, (k, PU.rel(o[k], r[k]))

    torch.manual_seed(3)
    enc = PointnetEncoder(input_feature_dim=4).to(DEV).eval()
    enc_r = R.PointnetEncoder(input_feature_dim=4).to(DEV).eval()
    enc_r.load_state_dict(enc.state_dict(), strict=True)
    with torch.no_grad():
        eo, er = enc(dict(data)), enc_r(dict(data))
    assert PU.rel(eo["enc_features"], er["enc_features"]) < RTOL and PU.rel(eo["enc_preds"], er["enc_preds"]) < RTOL
    # whole_scene=True: (B, num_bboxes, N, 3+C) objects with a validity mask -> same rows as encoding them directly
    enc.whole_scene = True
    objs = torch.from_numpy(pc).to(DEV).view(2, 2, 8000, -1)
    masks = torch.tensor([[1, 0], [1, 1]], device=DEV)
    with torch.no_grad():
        ws = enc({"point_clouds": objs, "target_masks": masks})
    flat = eo["enc_features"].view(2, 2, -1)
    assert PU.rel(ws["enc_features"][1], flat[1]) < 1e-5 and float(ws["enc_features"][0, 1].abs().max()) == 0.0
