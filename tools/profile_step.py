"""Kernel-level breakdown of one bench step with torch.profiler (CUPTI): top kernels by device time."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from torch.profiler import profile, ProfilerActivity

impl = sys.argv[1] if len(sys.argv) > 1 else "ours"
cfg = sys.argv[2] if len(sys.argv) > 2 else "c3"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
if impl == "reference":
    from conftest import load_reference_ext
    from oracle import ref_model as R
    R.set_backend(load_reference_ext())
host, num_words, C = bench.build_inputs(cfg, 42, 0)
model, DC, loss_fn = bench.build_model(impl, C, dev)
from scan2cap_b200.distributed import FlatGradients
flat = FlatGradients(model)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)
data = bench.to_device(host, dev, num_words)
def step():
    flat.zero_()
    out = loss_fn(model({k: v for k, v in data.items()}), dev, DC, None, **bench.LOSS_FLAGS)
    out["loss"].backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
import time
t = time.perf_counter(); step(); torch.cuda.synchronize(); print("wall ms/step (no profiler):", 1e3 * (time.perf_counter() - t))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ka = prof.key_averages()
rows = [(e.key, e.count, e.self_device_time_total) for e in ka if e.self_device_time_total > 0]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print("total device time (us): %.0f over %d kernel launches" % (tot, sum(r[1] for r in rows)))
for k, c, t in rows[:45]:
    print("%8.0f us %5.1f%% x%-5d %s" % (t, 100 * t / tot, c, k[:110]))
