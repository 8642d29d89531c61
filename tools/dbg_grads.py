import copy, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_engine_gpu as T
from scan2cap_b200.engine import TrainStep
model, DC, batches = T._setup()
me, mg = copy.deepcopy(model), copy.deepcopy(model)
junk = [torch.full((1 << 26,), float("nan"), device="cuda") for _ in range(8)]
junk += [torch.full((n,), float("nan"), device="cuda") for n in (1 << 20, 1 << 16, 1 << 12, 1 << 18, 1 << 22, 1 << 14, 1 << 10)] * 8
torch.cuda.synchronize(); del junk
eager = TrainStep(me, DC, use_cuda_graph=False, **T.FLAGS)
graph = TrainStep(mg, DC, use_cuda_graph=True, **T.FLAGS)
le = eager.run(dict(batches[0])); lg = graph.run(dict(batches[0]))
torch.cuda.synchronize()
print("loss", float(le), float(lg))
ge = {n: p.grad.clone() for n, p in me.named_parameters()}
gg = {n: p.grad.clone() for n, p in mg.named_parameters()}
rows = []
for n in ge:
    a, b = ge[n].double(), gg[n].double()
    rows.append((float((a - b).norm() / (a.norm() + 1e-30)), n, float(a.norm()), bool(torch.isnan(b).any())))
rows.sort(reverse=True)
for r in rows[:14]:
    print("%.3e %s |g|=%.3e nan=%s" % r)
oe, og = eager.last, graph.last
for k in ("loss", "vote_loss", "objectness_loss", "box_loss", "sem_cls_loss", "cap_loss", "ori_loss"):
    print(k, float(oe[k]), float(og[k]))
