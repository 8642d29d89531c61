"""One eager training step bracketed by cudaProfilerStart/Stop, for ncu --profile-from-start off."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
host, num_words, C = bench.build_inputs(cfg, 42, 0)
model, DC, loss_fn = bench.build_model("ours", C, dev)
from scan2cap_b200.engine import TrainStep
eng = TrainStep(model, DC, use_cuda_graph=False, **bench.LOSS_FLAGS)
data = bench.to_device(host, dev, num_words)
for _ in range(2):
    eng.run_eager(dict(data))
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.run_eager(dict(data))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
