"""Which source lines issue the framework (ATen) operators of one eager training step?
A TorchDispatchMode logs every non-view ATen op with the innermost scan2cap_b200 frame on the Python stack (forward, loss
and the Python backward of the custom autograd Functions; the backward of built-in ops runs on the autograd thread
without Python frames and is listed under its op name only).
usage: python tools/op_origins.py [c3|c4] > gpurun_out/op_origins.txt"""
import collections
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.utils._python_dispatch import TorchDispatchMode

import bench

VIEW = ("view", "expand", "slice", "select", "transpose", "unsqueeze", "squeeze", "reshape", "_unsafe_view", "detach",
        "alias", "permute", "as_strided", "empty", "t.default", "unbind", "split", "chunk", "narrow", "lift_fresh",
        "_local_scalar_dense", "is_", "size", "stride", "numel", "sym_", "unfold", "diagonal", "real", "view_as", "_to_copy_view",
        "new_empty", "empty_like", "empty_strided", "record_stream", "result_type", "can_cast")

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

rows = collections.Counter()


class Log(TorchDispatchMode):
    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = str(func).replace("aten.", "")
        base = name.split(".")[0]
        if not any(base == v or base.startswith(v) for v in VIEW):
            where = "(autograd thread / no package frame)"
            for fr in reversed(traceback.extract_stack(limit=40)):
                if "scan2cap_b200" in fr.filename and not fr.filename.endswith("_lib.py"):
                    where = "%s:%d %s" % (fr.filename.split("scan2cap_b200/")[-1], fr.lineno, fr.name)
                    break
            rows[(where, name)] += 1
        return func(*args, **(kwargs or {}))


with Log():
    o.engine.run_eager(dict(data))
torch.cuda.synchronize()
by_where = collections.defaultdict(list)
for (where, name), n in rows.items():
    by_where[where].append((n, name))
print("# non-view ATen ops of one eager %s step by issuing source line: %d ops" % (cfg, sum(rows.values())))
for where, ops in sorted(by_where.items(), key=lambda kv: -sum(n for n, _ in kv[1])):
    print("%4d  %s" % (sum(n for n, _ in ops), where))
    print("        " + ", ".join("%dx %s" % (n, nm) for n, nm in sorted(ops, reverse=True)))
