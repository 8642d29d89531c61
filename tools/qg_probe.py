"""One fused query+group launch at a sweep point (for ncu): python tools/qg_probe.py N C ns [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from scan2cap_b200 import synthetic
from scan2cap_b200.lib.pointnet2 import _ext
N, C, ns = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
B = int(sys.argv[4]) if len(sys.argv) > 4 else 8
pc, _ = synthetic.make_point_clouds(B, N, use_normal=False, use_height=False, seed=42)
full = torch.cat([torch.from_numpy(pc[..., :3].copy()), torch.randn(B, N, C)], -1).cuda()   # like point_clouds (B,N,3+C)
xyz = full[..., :3].contiguous()
feats = full[..., 3:]
if len(sys.argv) > 5 and sys.argv[5] == "aligned":  # (B,N,C) contiguous rows (16-byte aligned), as between SA levels
    feats = feats.contiguous()
r = 0.2 * (40000.0 / N) ** 0.5
_, new_xyz = _ext.furthest_point_sampling_with_xyz(xyz, 2048)
for _ in range(2):
    _ext.query_and_group(xyz, new_xyz, feats, r, ns, True, feat_point_major=True, channels_last=True, pad4=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
_ext.query_and_group(xyz, new_xyz, feats, r, ns, True, feat_point_major=True, channels_last=True, pad4=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
