import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
import test_capnet_gpu as T
from oracle import ref_loss as RL
from scan2cap_b200.lib.loss_helper import get_scene_cap_loss
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
try:
    print("fp32_precision", torch.backends.fp32_precision, torch.backends.cudnn.conv.fp32_precision, torch.backends.cuda.matmul.fp32_precision)
except Exception as e:
    print("no fp32_precision api", e)
DEV = "cuda:0"
V = 150
ours, ref, DC = T._models("center", 4, V)
data = T._data(2, 8000, V, seed=11)
with torch.no_grad():
    state = copy.deepcopy(ours.state_dict())
    probe = ours(T._clone(data))
    ours.load_state_dict(state)
data["ref_box_corner_label"] = probe["bbox_corner"][:, 7].clone()
data["ref_box_corner_label"][-1] += 50.0
ours.train(); ref.train()
WATCH = ["sa1_features", "sa2_features", "fp2_features", "vote_features", "vote_xyz", "aggregated_vote_features", "objectness_scores", "center", "size_scores", "sem_cls_scores", "lang_cap", "edge_orientations"]
def run(model, lossf, flags):
    out = model(T._clone(data))
    for k in WATCH:
        if out[k].requires_grad: out[k].retain_grad()
    out = lossf(out, DEV, DC, None, *flags)
    model.zero_grad()
    out["loss"].backward()
    return out
for flags in [(True, True, True, True), (True, False, False, False), (False, True, False, False)]:
    print("=== flags (det, cap, ori, dist):", flags)
    o = run(ours, get_scene_cap_loss, flags)
    torch.backends.cudnn.enabled = False
    r = run(ref, RL.get_scene_cap_loss, flags)
    torch.backends.cudnn.enabled = True
    for k in T.FLOAT_KEYS:
        e = T._rel(o[k], r[k])
        if e > 1e-5: print("  fwd %-28s %.2e" % (k, e))
    for k in WATCH:
        if o[k].grad is not None and r[k].grad is not None:
            print("  dL/d%-26s %.2e   |g|max %.3e" % (k, T._rel(o[k].grad, r[k].grad), float(r[k].grad.abs().max())))
    go = {n: p.grad for n, p in ours.named_parameters() if p.grad is not None}
    gr = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
    gmax = max(float(g.abs().max()) for g in gr.values())
    worst = sorted(((float((go[n]-gr[n]).abs().max())/max(float(gr[n].abs().max()),1e-3*gmax), n) for n in gr), reverse=True)[:6]
    print("  worst param grads:", worst)
# self-consistency of the oracle (atomics order) and of ours
r1 = run(ref, RL.get_scene_cap_loss, (True, True, True, True)); g1 = {n: p.grad.clone() for n, p in ref.named_parameters() if p.grad is not None}
r2 = run(ref, RL.get_scene_cap_loss, (True, True, True, True)); g2 = {n: p.grad.clone() for n, p in ref.named_parameters() if p.grad is not None}
print("oracle run-to-run worst grad rel:", max(T._rel(g1[n], g2[n]) for n in g1))
