"""Host->device bandwidth of the box (pinned memory) and the duration of TrainStep.prefetch's copy-stream work at c4.
usage: python tools/h2d_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
h = torch.empty(93_744_960 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(h, device=dev)
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); d.copy_(h, non_blocking=True); b.record(); b.synchronize()
    ts.append(a.elapsed_time(b))
print("pinned H2D of 93.7 MB: %.2f ms = %.1f GB/s" % (min(ts), 93.74496 / min(ts)))

o = bench.Ours("c4", dev, 0)
timer = bench.Timer(dev, 1)
o.timed(timer, 3, True)
eng = o.engine
cs = eng._copy_stream
for _ in range(3):
    nxt = dict(o.host, num_words=o.num_words)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(cs)
    eng.prefetch(nxt)
    b.record(cs)
    b.synchronize()
    print("prefetch (H2D of all inputs + FPS of all levels + SA1 grid) on an idle GPU: %.2f ms" % a.elapsed_time(b))
    eng.run(nxt)
torch.cuda.synchronize()
