"""Fused query+group (SA1 shape: B=8, N=40 000, M=2048, ns=64, C=132) -- time per ring geometry of the TMA epilogue
vs the LDG/STG epilogue, aligned (B,N,C) features and the padded point_clouds layout (row stride 136)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scan2cap_b200._lib as L  # noqa: E402
from scan2cap_b200 import synthetic  # noqa: E402
from scan2cap_b200.lib.pointnet2 import _ext  # noqa: E402

B, N, M, C = 8, 40000, 2048, 132
pc, _ = synthetic.make_point_clouds(B, N, use_normal=False, use_height=False, seed=42)
xyz = torch.from_numpy(np.ascontiguousarray(pc[..., :3])).cuda()
_, new_xyz = _ext.furthest_point_sampling_with_xyz(xyz, M)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
peak = 6553.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, iters=15):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


for ns in (64, 32, 16):
    alg = B * (12 * N + 12 * M + 4 * C * N + 4 * M * ns + 4 * (3 + C) * M * ns)
    for layout in ("aligned", "padded_pc"):
        if layout == "aligned":
            feats = torch.randn(B, N, C, device="cuda")
        else:
            buf = torch.randn(B, N, C + 4, device="cuda")
            feats = buf[..., 4:]
        ref = None
        for variant in (-1, 0, 1, 2, 3, 4, 5):
            L.LIB.s2c_query_and_group_grid_tune(variant)
            fn = lambda: _ext.query_and_group(xyz, new_xyz, feats, 0.2, ns, True, feat_point_major=True,
                                              channels_last=True, pad4=True)
            out = fn()[0]
            if ref is None:
                ref = out.clone()
            ok = bool(torch.equal(out, ref))
            med, mn = timeit(fn)
            print(json.dumps({"ns": ns, "layout": layout, "variant": variant, "ms_median": med, "ms_min": mn,
                              "frac_hbm": alg / med / 1e6 / peak, "equal_to_ldg": ok}), flush=True)
        del feats
L.LIB.s2c_query_and_group_grid_tune(0)
