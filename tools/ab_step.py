"""Same-box A/B of training-step variants (box-to-box noise of the c3 step is ~1 %, as large as the effects compared):
each variant gets its own model + captured graph in this process; the timed blocks alternate between the variants.
usage: python tools/ab_step.py [c3|c4] > gpurun_out/ab_step.txt"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.nn.functional as F
import bench

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
from scan2cap_b200.models import caption_module
from scan2cap_b200.lib import linear_simt
from scan2cap_b200 import distributed

variants = {}
variants["current"] = bench.Ours(cfg, dev, 0)
timer = bench.Timer(dev, 1)
variants["current"].timed(timer, 3, False)          # capture

caption_module.linear = lambda x, w, b=None: F.linear(x, w, b)
variants["library GEMMs in the caption module"] = bench.Ours(cfg, dev, 0)
variants["library GEMMs in the caption module"].timed(timer, 3, False)
caption_module.linear = linear_simt.linear

os.environ["S2C_MLP_ATM"] = "0"
variants["forward A operand through shared memory"] = bench.Ours(cfg, dev, 0)
variants["forward A operand through shared memory"].timed(timer, 3, False)
del os.environ["S2C_MLP_ATM"]

rel, col = distributed.FlatGradients.release, distributed.FlatGradients.collect
distributed.FlatGradients.release = lambda self: self.zero_()
distributed.FlatGradients.collect = lambda self: None
variants["per-parameter gradient accumulation"] = bench.Ours(cfg, dev, 0)
variants["per-parameter gradient accumulation"].timed(timer, 3, False)
distributed.FlatGradients.release, distributed.FlatGradients.collect = rel, col

res = {k: [] for k in variants}
for rep in range(4):
    for k, o in variants.items():
        ms, _ = o.timed(timer, 10, False)
        res[k].append(ms / 10)
print("# %s step, ms (4 alternating blocks of 10 graph replays each, L2 flushed between steps)" % cfg)
for k, v in res.items():
    print("%-40s median %.3f   blocks %s" % (k, float(np.median(v)), " ".join("%.3f" % x for x in v)))
