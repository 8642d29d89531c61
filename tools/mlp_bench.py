import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, numpy as np
import torch.nn.functional as F
from scan2cap_b200.lib.pointnet2 import _ext_mlp
torch.backends.cuda.matmul.allow_tf32 = False
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
for (R, K, N) in [(1048576, 8, 64), (1048576, 64, 64), (1048576, 64, 128), (262144, 132, 128), (262144, 128, 128), (262144, 128, 256), (65536, 260, 128), (32768, 260, 128), (8192, 512, 256)]:
    A = torch.randn(R, K, device="cuda"); W = torch.randn(N, K, device="cuda")
    sc = torch.rand(K, device="cuda") + 0.5; sh = torch.randn(K, device="cuda")
    t_k = timeit(lambda: _ext_mlp.mlp_layer_fwd(A, W, sc, sh, want_stats=True))
    def torch_path():
        x = torch.relu_(torch.addcmul(sh, A, sc)); y = F.linear(x, W); return y, y.sum(0), (y * y).sum(0)
    t_t = timeit(torch_path)
    t_mm = timeit(lambda: F.linear(A, W))
    bytes_ = 4.0 * R * (K + N)
    print(json.dumps(dict(R=R, K=K, N=N, ours_ms=t_k, torch_ms=t_t, linear_only_ms=t_mm, GBps=bytes_ / t_k / 1e6, frac_hbm=bytes_ / t_k / 1e6 / 6556.5)), flush=True)
