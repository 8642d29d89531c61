"""Timeline of the layer kernel's pipeline roles on CTA 0 (s2c_mlp_probe: clock64 stamps at the hand-offs).
Prints, per K chunk, when (in ns after kernel start, at the SM clock given) the loader issued its TMA loads, the
transform warps saw the raw tile / a free operand stage / finished staging, the MMA warp saw its operands and issued, and
per tile when the epilogue saw the accumulator and finished.
usage: python tools/mlp_pipe_probe.py > gpurun_out/mlp_pipe_probe.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from scan2cap_b200 import _lib
from scan2cap_b200.lib.pointnet2 import _ext_mlp

MHZ = 1965.0
CAP = 16 + 16 * 64
buf = torch.zeros(CAP, dtype=torch.int64, device="cuda")


def show(title, fn, KC, chunks=16):
    fn(); fn()
    torch.cuda.synchronize()
    buf.zero_()
    _lib.call("s2c_mlp_probe", buf.data_ptr(), CAP)
    fn()
    torch.cuda.synchronize()
    _lib.call("s2c_mlp_probe", None, 0)
    v = buf.cpu().numpy().astype("int64")
    t0 = v[0]
    ns = lambda x: (x - t0) / MHZ * 1e3 if x > 0 else float("nan")
    print("== %s   (KC = %d chunks per tile)" % (title, KC))
    print("chunk  rawTMA    wTMA | T:wait  rawOK   opFree  staged | M:ready  issued | tile: accFull  epiDone")
    for i in range(chunks):
        b = 16 + 16 * i
        r = v[b:b + 16]
        if r[0] == 0 and r[4] == 0:
            break
        print("%5d %7.0f %7.0f | %6.0f %6.0f %7.0f %7.0f | %7.0f %7.0f | %13.0f %8.0f" % (
            i, ns(r[6]), ns(r[7]), ns(r[0]), ns(r[1]), ns(r[2]), ns(r[3]), ns(r[4]), ns(r[5]), ns(r[8]), ns(r[9])))


for R, K, N in [(1048576, 64, 64), (1048576, 64, 128), (262144, 128, 128), (8192, 256, 128), (2048, 128, 128)]:
    A = torch.randn(R, K, device="cuda"); W = torch.randn(N, K, device="cuda")
    sc = torch.rand(K, device="cuda") + 0.5; sh = torch.randn(K, device="cuda")
    show("fwd R=%d %d->%d" % (R, K, N), lambda: _ext_mlp.mlp_layer_fwd(A, W, sc, sh, want_stats=True), (K + 31) // 32,
         chunks=24 if R > 100000 else 16)
for R, K, N in [(1048576, 64, 64), (262144, 128, 128)]:
    Y = torch.randn(R, K, device="cuda"); G = torch.randn(R, K, device="cuda"); Yp = torch.randn(R, N, device="cuda")
    W = torch.randn(K, N, device="cuda")
    a = torch.rand(K, device="cuda"); b = torch.randn(K, device="cuda") * 0.01; c = torch.randn(K, device="cuda") * 0.01
    sc = torch.rand(N, device="cuda") + 0.5; sh = torch.randn(N, device="cuda")
    show("bwd dense R=%d %d->%d" % (R, K, N), lambda: _ext_mlp.mlp_layer_bwd_data(Y, a, b, c, W, Yp, sc, sh, G=G),
         (K + 31) // 32, chunks=24)
