"""Ring-depth sweep of the grouped-MLP layer kernel (mlp2.cu): A-operand stages OS, weight stages WS (0 = resident),
raw stages RS, through the S2C_MLP_OS / S2C_MLP_WS / S2C_MLP_RS knobs of choose_pipe().  Forward and dense backward-data
at the shapes of the CapNet stacks.  usage: python tools/mlp_pipe_sweep.py > gpurun_out/mlp_pipe_sweep.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from scan2cap_b200.lib.pointnet2 import _ext_mlp

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=7):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.fill_(1)  # L2 flush between iterations
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)) * 1e3


def setenv(cfg):
    for k in ("S2C_MLP_OS", "S2C_MLP_WS", "S2C_MLP_RS"):
        os.environ.pop(k, None)
    if cfg is not None:
        os.environ["S2C_MLP_OS"], os.environ["S2C_MLP_WS"], os.environ["S2C_MLP_RS"] = [str(c) for c in cfg]


CFGS = [None, (2, 2, 4), (3, 0, 4), (2, 3, 2)]  # (the knobs act on the shared-memory operand path: backward kernels, or forward with S2C_MLP_ATM=0)
FWD = [(1048576, 8, 64), (1048576, 64, 64), (1048576, 64, 128), (262144, 132, 128), (262144, 128, 128), (65536, 260, 128),
       (65536, 128, 128), (32768, 128, 128), (8192, 256, 128), (20480, 256, 128), (2048, 128, 128)]
print("# us per launch (median of 7, L2 flushed); columns = (OS, WS, RS) request, 0 = resident weights; 'auto' = choose_pipe()")
print("%-28s" % "shape" + "".join("%11s" % ("auto" if c is None else "%d/%d/%d" % c) for c in CFGS))
for R, K, N in FWD:
    A = torch.randn(R, K, device="cuda"); W = torch.randn(N, K, device="cuda")
    sc = torch.rand(K, device="cuda") + 0.5; sh = torch.randn(K, device="cuda")
    row = []
    for cfg in CFGS:
        setenv(cfg)
        row.append(timeit(lambda: _ext_mlp.mlp_layer_fwd(A, W, sc, sh, want_stats=True)))
    floor = 4.0 * R * (K + N) / 6553e3
    print("%-28s" % ("fwd R=%d %d->%d" % (R, K, N)) + "".join("%11.1f" % t for t in row) + "   floor %.1f" % floor, flush=True)
    del A, W
BWD = [(1048576, 128, 64), (1048576, 64, 64), (262144, 128, 128), (262144, 256, 128), (65536, 128, 128), (32768, 256, 128),
       (20480, 128, 128), (8192, 256, 256)]
for R, K, N in BWD:  # layer l has K channels, layer l-1 has N
    Y = torch.randn(R, K, device="cuda"); G = torch.randn(R, K, device="cuda"); Yp = torch.randn(R, N, device="cuda")
    W = torch.randn(K, N, device="cuda")
    a = torch.rand(K, device="cuda"); b = torch.randn(K, device="cuda") * 0.01; c = torch.randn(K, device="cuda") * 0.01
    sc = torch.rand(N, device="cuda") + 0.5; sh = torch.randn(N, device="cuda")
    row = []
    for cfg in CFGS:
        setenv(cfg)
        row.append(timeit(lambda: _ext_mlp.mlp_layer_bwd_data(Y, a, b, c, W, Yp, sc, sh, G=G)))
    floor = 4.0 * R * (3 * K + 2 * N) / 6553e3
    print("%-28s" % ("bwd R=%d %d->%d" % (R, K, N)) + "".join("%11.1f" % t for t in row) + "   floor %.1f" % floor, flush=True)
    del Y, G, Yp, W
# pooled-gradient variant (last layer of an SA stack)
for R, K, N, ns in [(1048576, 128, 64, 64), (262144, 256, 128, 32), (65536, 256, 128, 16)]:
    Y = torch.randn(R, K, device="cuda"); Yp = torch.randn(R, N, device="cuda"); W = torch.randn(K, N, device="cuda")
    Gp = R // ns
    dpool = torch.randn(Gp, K, device="cuda"); am = torch.randint(0, ns, (Gp, K), device="cuda", dtype=torch.int32)
    a = torch.rand(K, device="cuda"); b = torch.randn(K, device="cuda") * 0.01; c = torch.randn(K, device="cuda") * 0.01
    sc = torch.rand(N, device="cuda") + 0.5; sh = torch.randn(N, device="cuda")
    ls = torch.rand(K, device="cuda") + 0.5; lh = torch.randn(K, device="cuda")
    row = []
    for cfg in CFGS:
        setenv(cfg)
        row.append(timeit(lambda: _ext_mlp.mlp_layer_bwd_data(Y, a, b, c, W, Yp, sc, sh, dpool=dpool, argmax=am, ns=ns,
                                                               last_scale=ls, last_shift=lh)))
    floor = 4.0 * R * (2 * K + 2 * N) / 6553e3
    print("%-28s" % ("pool R=%d %d->%d ns%d" % (R, K, N, ns)) + "".join("%11.1f" % t for t in row) + "   floor %.1f" % floor,
          flush=True)
    del Y, Yp, W
