"""The caption module's Linear layers: libs2c's fp32 GEMM kernels (lib/linear_simt.py) vs the framework's cuBLAS calls,
piece by piece (forward, input gradient, weight gradient).  usage: python tools/gemm_bench.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from scan2cap_b200._lib import call
from scan2cap_b200.lib.linear_simt import gemm, _stream
torch.backends.cuda.matmul.allow_tf32 = False


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)) * 1e3


print("%-34s %10s %10s" % ("piece (R, K -> N)", "ours us", "cuBLAS us"))
for name, R, K, N in [("map_feat", 2048, 128, 512), ("pre_word", 216, 300, 300), ("pre_tgt", 8, 128, 300),
                      ("classifier", 216, 512, 3500), ("xyz grad", 32768, 128, 4)]:
    x = torch.randn(R, K, device="cuda"); w = torch.randn(N, K, device="cuda"); dy = torch.randn(R, N, device="cuda")
    dw = torch.empty(N, K, device="cuda"); db = torch.empty(N, device="cuda")
    f = timeit(lambda: gemm(x, (w.stride(1), w.stride(0)), w, R, N, K))
    fb = timeit(lambda: torch.nn.functional.linear(x, w))
    d = timeit(lambda: gemm(dy, (w.stride(0), w.stride(1)), w, R, K, N))
    dbl = timeit(lambda: dy @ w)
    g = timeit(lambda: call("s2c_gemm_tn", dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), R, N, K, dw.data_ptr(), K,
                            db.data_ptr(), _stream(x)))
    gb = timeit(lambda: dy.t() @ x)
    print("%-34s %10.1f %10.1f" % ("%s fwd (%d, %d -> %d)" % (name, R, K, N), f, fb))
    print("%-34s %10.1f %10.1f" % ("%s dX" % name, d, dbl))
    print("%-34s %10.1f %10.1f" % ("%s dW (+db)" % name, g, gb))
