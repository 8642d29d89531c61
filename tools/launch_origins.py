"""Where do the framework (ATen / cuBLAS) launches of one eager training step come from?
Runs one step under torch.profiler with Python stacks and attributes every framework kernel to the first
scan2cap_b200 source line on its stack (forward) or to the autograd node + the forward line of the same
sequence number (backward).  usage: python tools/launch_origins.py [c3|c4] > gpurun_out/launch_origins.txt"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

import bench

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
o = bench.Ours(cfg, dev, 0, use_graph=False)
data = o.resident()
for _ in range(2):
    o.engine.run_eager(dict(data))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    o.engine.run_eager(dict(data))
    torch.cuda.synchronize()


def frame(stack):
    for s in stack or []:
        if "scan2cap_b200" in s and "_lib.py" not in s:
            return s.split("scan2cap_b200/")[-1]
    return None


events = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU]
fwd_frame = {}
for e in events:
    if e.sequence_nr is not None and e.sequence_nr >= 0 and e.stack:
        f = frame(e.stack)
        if f and e.sequence_nr not in fwd_frame:
            fwd_frame[e.sequence_nr] = f

rows = collections.defaultdict(lambda: [0, 0.0, collections.Counter()])
total = 0
for e in events:
    if not e.kernels:
        continue
    f = frame(e.stack)
    if f is None:
        p, node = e, None
        while p is not None:
            if p.name.startswith("autograd::engine::evaluate_function"):
                node = p
            p = p.cpu_parent
        if node is not None:
            f = "BWD %s <- %s" % (node.name.split(": ")[-1], fwd_frame.get(node.sequence_nr, "?"))
        else:
            f = "(no frame) " + e.name
    for k in e.kernels:
        r = rows[f]
        r[0] += 1
        r[1] += k.duration
        r[2][k.name[:60]] += 1
        total += 1

print("# framework-launched kernels of one eager %s step by origin: %d launches" % (cfg, total))
for f, (n, us, names) in sorted(rows.items(), key=lambda kv: -kv[1][0]):
    print("%4d %8.1f us  %s" % (n, us, f))
    for nm, c in names.most_common(4):
        print("            %3d x %s" % (c, nm))
