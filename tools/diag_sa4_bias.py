import os, sys, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_capnet_gpu as T
from oracle import ref_loss as RL
from scan2cap_b200.lib.loss_helper import get_scene_cap_loss
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
V = 150
ours, ref, DC = T._models("corner", 4, V)
data = T._data(1, 20000, V, seed=11)
with torch.no_grad():
    state = copy.deepcopy(ours.state_dict()); probe = ours(T._clone(data)); ours.load_state_dict(state)
data["ref_box_corner_label"] = probe["bbox_corner"][:, 7].clone(); data["ref_box_corner_label"][-1] += 50.0
ours.train(); ref.train()
o = get_scene_cap_loss(ours(T._clone(data)), "cuda:0", DC, None, True, True, True, True)
o["sa4_features"].retain_grad(); o["sa3_features"].retain_grad()
with torch.backends.cudnn.flags(enabled=False):
    r = RL.get_scene_cap_loss(ref(T._clone(data)), "cuda:0", DC, None, True, True, True, True)
    r["sa4_features"].retain_grad(); r["sa3_features"].retain_grad()
    r["loss"].backward()
o["loss"].backward()
l2 = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
for name in ("sa4", "sa3"):
    fo, fr = o[name + "_features"], r[name + "_features"]        # (B, C, M)
    go, gr = fo.grad, fr.grad
    print(name, "features L2", l2(fo, fr), " dL/dfeatures L2", l2(go, gr))
    chk_o = (go * (fo > 0)).sum((0, 2)); chk_r = (gr * (fr > 0)).sum((0, 2))
    po = dict(ours.named_parameters())["backbone_net.%s.mlp_module.layer2.bn.bn.bias" % name].grad
    pr = dict(ref.named_parameters())["backbone_net.%s.mlp_module.layer2.bn.bn.bias" % name].grad
    print("   ours: bias.grad vs sum(dpool*[pooled>0])  L2 %.2e ;  oracle: same check L2 %.2e ; ours-vs-oracle bias.grad L2 %.2e ; check_o vs check_r %.2e" % (l2(po, chk_o), l2(pr, chk_r), l2(po, pr), l2(chk_o, chk_r)))
    print("   frac pooled==0: ours %.4f oracle %.4f ; |bias.grad| max ours %.3e oracle %.3e ; sum|dpool| %.3e" % (float((fo == 0).float().mean()), float((fr == 0).float().mean()), float(po.abs().max()), float(pr.abs().max()), float(gr.abs().sum())))
