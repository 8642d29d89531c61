"""Times of the fused query+group entry (SA1 / C=132 / ns=64) for a few variants, plus ball_query alone."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scan2cap_b200._lib as L
from scan2cap_b200 import synthetic
from scan2cap_b200.lib.pointnet2 import _ext
B, N, M, C = 8, 40000, 2048, 132
pc, _ = synthetic.make_point_clouds(B, N, use_normal=False, use_height=False, seed=42)
xyz = torch.from_numpy(np.ascontiguousarray(pc[..., :3])).cuda()
_, new_xyz = _ext.furthest_point_sampling_with_xyz(xyz, M)
feats = torch.randn(B, N, C, device="cuda")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
def timeit(fn, iters=15):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    return round(float(np.median(ts)) * 1e3, 1)
print("ball_query alone (build + query) us:", timeit(lambda: _ext.ball_query(new_xyz, xyz, 0.2, 64)))
for ns in (64, 16):
    for v in [int(x) for x in os.environ.get("QG_VARIANTS", "-1,0,1,2").split(",")]:
        L.LIB.s2c_query_and_group_grid_tune(v)
        t = timeit(lambda: _ext.query_and_group(xyz, new_xyz, feats, 0.2, ns, True, feat_point_major=True, channels_last=True, pad4=True))
        alg = B * (12 * N + 12 * M + 4 * C * N + 4 * M * ns + 4 * (3 + C) * M * ns)
        print("ns", ns, "variant", v, "us", t, "frac", round(alg / t / 1e3 / 6553.0, 4), flush=True)
