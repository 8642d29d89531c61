import os, sys, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_capnet_gpu as T
from oracle import ref_loss as RL
from scan2cap_b200.lib.loss_helper import get_scene_cap_loss
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
qm, B, N = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
V = 150
ours, ref, DC = T._models(qm, 4, V)
data = T._data(B, N, V, seed=11)
with torch.no_grad():
    state = copy.deepcopy(ours.state_dict()); probe = ours(T._clone(data)); ours.load_state_dict(state)
data["ref_box_corner_label"] = probe["bbox_corner"][:, 7].clone(); data["ref_box_corner_label"][-1] += 50.0
ours.train(); ref.train()
o = get_scene_cap_loss(ours(T._clone(data)), "cuda:0", DC, None, True, True, True, True)
with torch.backends.cudnn.flags(enabled=False):
    r = RL.get_scene_cap_loss(ref(T._clone(data)), "cuda:0", DC, None, True, True, True, True)
    r["loss"].backward()
o["loss"].backward()
print("vote inds equal:", torch.equal(o["aggregated_vote_inds"], r["aggregated_vote_inds"]), "loss", float(o["loss"]), float(r["loss"]))
go = {n: p.grad for n, p in ours.named_parameters() if p.grad is not None}
gr = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
gmax = max(float(g.abs().max()) for g in gr.values())
rows = []
for n in gr:
    scale = max(float(gr[n].abs().max()), 1e-3 * gmax)
    e = float((go[n] - gr[n]).abs().max()) / scale
    l2 = float((go[n].double() - gr[n].double()).norm() / max(float(gr[n].double().norm()), 1e-3 * gmax))
    rows.append((e, l2, n, float(gr[n].abs().max())))
for e, l2, n, m in sorted(rows, reverse=True)[:14]:
    print("max-dev %.2e  L2 %.2e  |g|max %.2e  %s" % (e, l2, m, n))
