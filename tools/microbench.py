"""Per-op timings: libs2c (ours) vs the reference CUDA extension re-compiled for sm_100 (oracle/_ref),
same inputs, same GPU, CUDA events.  Usage:  python tools/microbench.py [--out gpurun_out/microbench.json]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_reference_ext  # noqa: E402
from scan2cap_b200 import synthetic  # noqa: E402
from scan2cap_b200.lib.pointnet2 import _ext  # noqa: E402


def timeit(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "microbench.json"))
    ap.add_argument("--batch", type=int, default=8)
    args = ap.parse_args()
    ref = load_reference_ext()
    B = args.batch
    rows = []

    def row(name, ours, theirs, **kw):
        r = dict(op=name, ours_ms=ours, ref_ms=theirs, speedup=(theirs / ours if theirs else None), **kw)
        rows.append(r)
        print(json.dumps(r), flush=True)

    pc, _ = synthetic.make_point_clouds(B, 40000, use_normal=True, seed=42)
    pcs = torch.from_numpy(pc).cuda()
    xyz = pcs[..., :3].contiguous()
    # FPS chain of the backbone
    cur = xyz
    for (n, m) in [(40000, 2048), (2048, 1024), (1024, 512), (512, 256)]:
        t_o = timeit(lambda: _ext.furthest_point_sampling(cur, m))
        t_r = timeit(lambda: ref.furthest_point_sampling(cur, m)) if ref else None
        row("fps", t_o, t_r, B=B, n=n, m=m, us_per_pick=1e3 * t_o / (m - 1))
        idx, cur = _ext.furthest_point_sampling_with_xyz(cur, m)
    # ball query / group at SA shapes
    feats_cm = pcs[..., 3:].transpose(1, 2).contiguous()
    C = feats_cm.shape[1]
    idx1, new_xyz = _ext.furthest_point_sampling_with_xyz(xyz, 2048)
    for (M, r, ns) in [(2048, 0.2, 64), (2048, 0.2, 32), (2048, 0.2, 16)]:
        t_o = timeit(lambda: _ext.ball_query(new_xyz, xyz, r, ns))
        t_r = timeit(lambda: ref.ball_query(new_xyz, xyz, r, ns)) if ref else None
        row("ball_query", t_o, t_r, B=B, n=40000, M=M, r=r, ns=ns)
        idx = _ext.ball_query(new_xyz, xyz, r, ns)
        t_o = timeit(lambda: _ext.group_points(feats_cm, idx))
        t_r = timeit(lambda: ref.group_points(feats_cm, idx)) if ref else None
        gb = B * (4 * C * 40000 + 4 * M * ns + 4 * C * M * ns) / 1e9
        row("group_points", t_o, t_r, B=B, C=C, M=M, ns=ns, GBps=gb / (t_o * 1e-3))
        for cl in (False, True):
            t_o = timeit(lambda: _ext.query_and_group(xyz, new_xyz, pcs[..., 3:], r, ns, True, feat_point_major=True,
                                                      channels_last=cl))

            def ref_qg():
                i = ref.ball_query(new_xyz, xyz, r, ns)
                g = ref.group_points(xyz.transpose(1, 2).contiguous(), i)
                g -= new_xyz.transpose(1, 2).unsqueeze(-1)
                g /= r
                return torch.cat([g, ref.group_points(feats_cm, i)], 1)
            t_r = timeit(ref_qg) if ref else None
            gb = B * (12 * 40000 + 12 * M + 4 * C * 40000 + 4 * M * ns + 4 * (3 + C) * M * ns) / 1e9
            row("query_and_group" + ("_cl" if cl else ""), t_o, t_r, B=B, C=C, M=M, ns=ns, alg_GB=gb,
                GBps=gb / (t_o * 1e-3), frac_hbm=gb / (t_o * 1e-3) / 6556.5)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
