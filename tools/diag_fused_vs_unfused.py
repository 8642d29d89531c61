import os, sys, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_capnet_gpu as T
from scan2cap_b200.lib.loss_helper import get_scene_cap_loss
import scan2cap_b200.lib.pointnet2.pointnet2_modules as pm
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
qm, B, N = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
V = 150
ours, ref, DC = T._models(qm, 4, V)
data = T._data(B, N, V, seed=11)
state = copy.deepcopy(ours.state_dict())
with torch.no_grad():
    probe = ours(T._clone(data)); ours.load_state_dict(state)
data["ref_box_corner_label"] = probe["bbox_corner"][:, 7].clone(); data["ref_box_corner_label"][-1] += 50.0
ours.train()
res = {}
for fused in (False, True):
    pm.USE_FUSED_MLP = fused
    ours.load_state_dict(state); ours.zero_grad()
    o = get_scene_cap_loss(ours(T._clone(data)), "cuda:0", DC, None, True, True, True, True)
    for k in ("sa4_features", "sa3_features", "fp2_features"): o[k].retain_grad()
    o["loss"].backward()
    res[fused] = ({n: p.grad.clone() for n, p in ours.named_parameters() if p.grad is not None}, {k: o[k].grad.clone() for k in ("sa4_features", "sa3_features", "fp2_features")}, float(o["loss"]), o["aggregated_vote_inds"].clone())
print("losses", res[False][2], res[True][2], "vote inds equal", torch.equal(res[False][3], res[True][3]))
for k in res[False][1]:
    a, b = res[True][1][k], res[False][1][k]
    print("dL/d%s  L2 %.2e" % (k, float((a - b).norm() / b.norm())))
gmax = max(float(g.abs().max()) for g in res[False][0].values())
rows = []
for n, gr in res[False][0].items():
    go = res[True][0][n]
    l2 = float((go.double() - gr.double()).norm() / max(float(gr.double().norm()), 1e-3 * gmax))
    rows.append((l2, n))
for l2, n in sorted(rows, reverse=True)[:12]:
    print("L2 %.2e  %s" % (l2, n))
