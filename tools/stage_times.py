"""Device time per stage of one training step (CUDA-graph-free, but measured with events so CPU launch gaps count
only where the GPU actually idles; each stage is bracketed by synchronize to isolate device time)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from torch.profiler import profile, ProfilerActivity
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
host, num_words, C = bench.build_inputs("c3", 42, 0)
model, DC, loss_fn = bench.build_model("ours", C, dev)
data = bench.to_device(host, dev, num_words)

def dev_time(fn):
    """sum of kernel durations (device busy time) of fn via the profiler"""
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        out = fn(); torch.cuda.synchronize()
    t = sum(e.self_device_time_total for e in prof.key_averages() if e.device_type.name == "CUDA" or True) / 2.0
    evs = [e for e in prof.events() if e.device_type.name == "CUDA"]
    t = sum(e.device_time for e in evs)
    return out, t / 1000.0, len(evs)

for it in range(2):
    d = dict(data)
    res = {}
    d, res["backbone_fwd"], n1 = dev_time(lambda: model.backbone_net(d))
    def vote():
        xyz, f = model.vgen(d["fp2_xyz"], d["fp2_features"])
        d["seed_inds"], d["seed_xyz"], d["seed_features"] = d["fp2_inds"], d["fp2_xyz"], d["fp2_features"]
        f = f.div(torch.norm(f, p=2, dim=1).unsqueeze(1))
        d["vote_xyz"], d["vote_features"] = xyz, f
        return model.proposal(xyz, f, d)
    d, res["vote_proposal_fwd"], n2 = dev_time(vote)
    d, res["graph_fwd"], n3 = dev_time(lambda: model.graph(d))
    d, res["caption_fwd"], n4 = dev_time(lambda: model.caption(d, True, False))
    d, res["loss_fwd"], n5 = dev_time(lambda: loss_fn(d, dev, DC, None, **bench.LOSS_FLAGS))
    model.zero_grad()
    _, res["backward_all"], n6 = dev_time(lambda: d["loss"].backward())
    if it == 1:
        print({k: round(v, 3) for k, v in res.items()}, "kernels", (n1, n2, n3, n4, n5, n6), "total ms", round(sum(res.values()), 2))
# backward split: run backward of sub-losses to attribute
