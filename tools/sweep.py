"""BASELINE.json configs[4]: FPS + ball_query + group_points sweep, N in 10k..200k, nsample in {16,32,64},
C in {1, 132}, M = 2048 centres, B = 8 scenes, 1xB200; achieved GB/s of the ALGORITHMIC bytes
against the HBM roofline, next to the reference's own lib/pointnet2 kernels (oracle/_ref, sm_100 build) on the
same inputs.  Rows whose byte floor is below launch latency / the serial FPS chain are flagged, not hidden.

    python tools/sweep.py [--out gpurun_out/sweep.json] [--quick]
"""
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


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                v = float(json.load(f)["hbm_gbs"])
            if v > 0:
                return v, "MEASURED_PEAKS.json"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


FLUSH = None


def timeit(fn, warmup=2, iters=7, flush=True):
    """median CUDA-event time in ms; a 256 MB buffer is rewritten before every timed call (L2 flush)."""
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            FLUSH.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--quick", action="store_true", help="N in {10k, 40k, 200k} only")
    ap.add_argument("--no-ref", action="store_true")
    args = ap.parse_args()
    ref = None if args.no_ref else load_reference_ext()
    peak, peak_src = hbm_peak()
    B, M = args.batch, 2048
    rows = []

    def row(**kw):
        rows.append(kw)
        print(json.dumps(kw), flush=True)

    Ns = [10000, 40000, 200000] if args.quick else [10000, 20000, 40000, 80000, 120000, 200000]
    for N in Ns:
        pc, _ = synthetic.make_point_clouds(B, N, use_normal=False, use_height=False, seed=42)
        xyz = torch.from_numpy(np.ascontiguousarray(pc[..., :3])).cuda()
        radius = 0.2 * (40000.0 / N) ** 0.5  # keeps the expected ball occupancy (surface density ~ N)
        t_o = timeit(lambda: _ext.furthest_point_sampling(xyz, M), iters=3, warmup=1)
        t_r = timeit(lambda: ref.furthest_point_sampling(xyz, M), iters=3, warmup=1) if ref else None
        row(op="fps", N=N, M=M, B=B, ours_ms=t_o, ref_ms=t_r, us_per_pick=1e3 * t_o / (M - 1),
            bound="latency (serial arg-max chain); bytes 12N+4M are irrelevant")
        _, new_xyz = _ext.furthest_point_sampling_with_xyz(xyz, M)
        for C in (1, 132):
            g = torch.Generator(device="cuda").manual_seed(N + C)
            feats_pm = torch.randn((B, N, C), generator=g, device="cuda")           # point-major (ours)
            feats_cm = feats_pm.transpose(1, 2).contiguous() if ref else None       # (B,C,N) (reference layout)
            for ns in (16, 32, 64):
                idx = _ext.ball_query(new_xyz, xyz, radius, ns)
                if C == 1:
                    t_o = timeit(lambda: _ext.ball_query(new_xyz, xyz, radius, ns))
                    t_r = timeit(lambda: ref.ball_query(new_xyz, xyz, radius, ns), iters=3, warmup=1) if ref else None
                    if ref:
                        assert torch.equal(idx, ref.ball_query(new_xyz, xyz, radius, ns)), "ball_query mismatch"
                    alg = B * (12 * N + 12 * M + 4 * M * ns)
                    row(op="ball_query", N=N, M=M, ns=ns, B=B, r=radius, ours_ms=t_o, ref_ms=t_r, alg_bytes=alg,
                        GBps=alg / t_o / 1e6, frac_hbm=alg / t_o / 1e6 / peak,
                        mean_ball_fill=float((idx != idx[..., :1]).sum(-1).float().mean().item() + 1) / ns)
                # group_points alone (reference op; ours on the reference's (B,C,N) layout)
                if ref:
                    t_o = timeit(lambda: _ext.group_points(feats_cm, idx))
                    t_r = timeit(lambda: ref.group_points(feats_cm, idx), iters=3, warmup=1)
                    alg = B * (4 * C * N + 4 * M * ns + 4 * C * M * ns)
                    row(op="group_points", N=N, M=M, ns=ns, C=C, B=B, ours_ms=t_o, ref_ms=t_r, alg_bytes=alg,
                        GBps=alg / t_o / 1e6, frac_hbm=alg / t_o / 1e6 / peak)
                # fused QueryAndGroup.forward
                # the product layout: channels-last rows [x,y,z,0 | features | pad] (what PointnetSAModuleVotes requests)
                t_o = timeit(lambda: _ext.query_and_group(xyz, new_xyz, feats_pm, radius, ns, True,
                                                          feat_point_major=True, channels_last=True, pad4=True))

                def ref_qg():
                    i = ref.ball_query(new_xyz, xyz, radius, ns)
                    gx = ref.group_points(xyz.transpose(1, 2).contiguous(), i)
                    gx -= new_xyz.transpose(1, 2).unsqueeze(-1)
                    gx /= radius
                    return torch.cat([gx, ref.group_points(feats_cm, i)], 1)
                t_r = timeit(ref_qg, iters=3, warmup=1) if ref else None
                alg = B * (12 * N + 12 * M + 4 * C * N + 4 * M * ns + 4 * (3 + C) * M * ns)
                floor_us = alg / peak / 1e3
                row(op="query_and_group", N=N, M=M, ns=ns, C=C, B=B, r=radius, ours_ms=t_o, ref_ms=t_r, alg_bytes=alg,
                    GBps=alg / t_o / 1e6, frac_hbm=alg / t_o / 1e6 / peak, byte_floor_us=floor_us,
                    note=("byte floor %.1f us is at launch-latency scale" % floor_us) if floor_us < 20 else "")
            del feats_pm, feats_cm
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"hbm_peak_GBps": peak, "peak_source": peak_src, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
