"""One eager training step bracketed by cudaProfilerStart/Stop, for ncu --profile-from-start off.
usage: ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python tools/ncu_step.py [c3|c4]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
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
torch.cuda.profiler.start()
o.engine.run_eager(dict(data))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
