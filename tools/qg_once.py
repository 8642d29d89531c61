"""One fused query+group call at the SA1 / C=132 shape (for ncu captures)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scan2cap_b200._lib as L
from scan2cap_b200 import synthetic
from scan2cap_b200.lib.pointnet2 import _ext
B, N, M, C, ns = 8, 40000, 2048, int(os.environ.get("QG_C", "132")), int(os.environ.get("QG_NS", "64"))
pc, _ = synthetic.make_point_clouds(B, N, use_normal=False, use_height=False, seed=42)
xyz = torch.from_numpy(np.ascontiguousarray(pc[..., :3])).cuda()
_, new_xyz = _ext.furthest_point_sampling_with_xyz(xyz, M)
feats = torch.randn(B, N, C, device="cuda")
L.LIB.s2c_query_and_group_grid_tune(int(os.environ.get("QG_VARIANT", "0")))
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
for _ in range(int(os.environ.get("QG_ITERS", "3"))):
    flush.zero_()
    _ext.query_and_group(xyz, new_xyz, feats, 0.2, ns, True, feat_point_major=True, channels_last=True, pad4=True)
torch.cuda.synchronize()
