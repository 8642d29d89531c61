"""Fused detection-loss kernel (csrc/loss.cu: forward + backward in one launch) vs the framework formulation of the same
reference arithmetic (lib/loss_helper.py:24-187 of the reference, mirrored in scan2cap_b200/lib/loss_helper.py) and vs
a float64 evaluation of it."""
import numpy as np
import pytest
import torch

from scan2cap_b200 import synthetic
from scan2cap_b200.data.scannet.model_util_scannet import ScannetDatasetConfig

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KEYS = ("vote_loss", "objectness_loss", "center_loss", "heading_cls_loss", "heading_reg_loss", "size_cls_loss",
        "size_reg_loss", "sem_cls_loss", "box_loss", "obj_acc", "pos_ratio", "neg_ratio", "loss")


def _inputs(B, seed):
    DC = ScannetDatasetConfig()
    g = torch.Generator().manual_seed(seed)
    N, S, K = 3000, 1024, 256
    d = synthetic.make_data_dict(B, N, use_normal=False, num_vocabs=30, seed=seed)
    data = {k: torch.from_numpy(v).to(DEV) for k, v in d.items()}
    sa1 = torch.stack([torch.randperm(N, generator=g)[:2048] for _ in range(B)]).int().to(DEV)
    data["seed_inds"] = sa1[:, :S]                                            # a strided int32 view, as in the model
    data["seed_xyz"] = torch.gather(data["point_clouds"][..., :3], 1, data["seed_inds"].long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    leaves = {}
    leaves["vote_xyz"] = (data["seed_xyz"] + 0.3 * torch.randn(B, S, 3, generator=g).to(DEV)).requires_grad_(True)
    # proposals near the GT boxes so that all three objectness zones (near / grey / far) are populated
    centers = data["center_label"][:, :12]
    pick = torch.randint(0, 12, (B, K), generator=g).to(DEV)
    agg = torch.gather(centers, 1, pick.unsqueeze(-1).expand(-1, -1, 3)) + 0.35 * torch.randn(B, K, 3, generator=g).to(DEV)
    agg[:, ::9] = 0.05 * torch.randn(B, (K + 8) // 9, 3, generator=g).to(DEV)   # proposals next to the padded GT slots at 0
    data["aggregated_vote_xyz"] = agg.contiguous()
    W = 2 + 3 + 2 * DC.num_heading_bin + 4 * DC.num_size_cluster + DC.num_class
    leaves["net"] = (0.5 * torch.randn(B, K, W, generator=g)).to(DEV).requires_grad_(True)
    return DC, data, leaves


def _decode(data, leaves, DC, dtype=torch.float32):
    """The slices ProposalModule.decode_scores hands to the loss (proposal_module.py:105-144)."""
    d = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in data.items()}
    net = leaves["net"].to(dtype)
    NH, NS = DC.num_heading_bin, DC.num_size_cluster
    d["vote_xyz"] = leaves["vote_xyz"].to(dtype)
    d["objectness_scores"] = net[:, :, 0:2]
    d["center"] = d["aggregated_vote_xyz"] + net[:, :, 2:5]
    d["heading_scores"] = net[:, :, 5:5 + NH]
    d["heading_residuals_normalized"] = net[:, :, 5 + NH:5 + 2 * NH]
    d["size_scores"] = net[:, :, 5 + 2 * NH:5 + 2 * NH + NS]
    d["size_residuals_normalized"] = net[:, :, 5 + 2 * NH + NS:5 + 2 * NH + 4 * NS].reshape(net.shape[0], net.shape[1], NS, 3)
    d["sem_cls_scores"] = net[:, :, 5 + 2 * NH + 4 * NS:]
    return d, net


@pytest.mark.parametrize("B,seed", [(2, 1), (8, 2)])
def test_fused_detection_loss_matches_framework_path(B, seed):
    from scan2cap_b200.lib.loss_helper import get_scene_cap_loss
    DC, data, leaves = _inputs(B, seed)
    outs, grads = {}, {}
    for mode in ("fused", "torch", "f64"):
        for t in leaves.values():
            t.grad = None
        d, net = _decode(data, leaves, DC, torch.float64 if mode == "f64" else torch.float32)
        if mode == "fused":
            d["_head_outputs"] = net
        o = get_scene_cap_loss(d, DEV, DC, None, detection=True, caption=False, orientation=False, distance=False)
        o["loss"].backward()
        outs[mode] = o
        grads[mode] = {k: t.grad.clone() for k, t in leaves.items()}
    f, t, r = outs["fused"], outs["torch"], outs["f64"]
    for k in ("objectness_label", "object_assignment"):
        assert torch.equal(f[k].long(), t[k].long()), k
    assert torch.equal(f["objectness_mask"], t["objectness_mask"])
    assert 0 < int(f["objectness_label"].sum()) < f["objectness_label"].numel()
    for k in KEYS:
        a, b, c = float(f[k]), float(t[k]), float(r[k])
        assert abs(a - b) <= 1e-5 * max(1.0, abs(b)), (k, a, b)
        assert abs(a - c) <= 1e-5 * max(1.0, abs(c)), (k, a, c)
    for k in leaves:
        ref64 = grads["f64"][k].double()
        scale = float(ref64.abs().max())
        assert float((grads["fused"][k].double() - ref64).abs().max()) <= 1e-5 * scale, k
        assert float((grads["torch"][k].double() - ref64).abs().max()) <= 1e-5 * scale, k
