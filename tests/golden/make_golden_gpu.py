"""Generate tests/golden/native_ops.npz by running the UNMODIFIED reference CUDA extension
(oracle/_ref/pointnet2_ref_ext.so, built from /root/reference/lib/pointnet2/_ext_src by
oracle/build_ref.py) on the seeded inputs of tests/golden_cases.py.  Needs a GPU:

    gpurun -- python tests/golden/make_golden_gpu.py        # writes gpurun_out/golden/native_ops.npz

The file is then copied to tests/golden/ and committed; only OUTPUTS are stored (inputs are re-made
from the seeds).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import golden_cases as gc  # noqa: E402
from conftest import load_reference_ext  # noqa: E402


def main():
    ref = load_reference_ext()
    assert ref is not None, "oracle/_ref/pointnet2_ref_ext.so missing"
    dev = torch.device("cuda:0")
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = {}
    for name, (xyz, m) in gc.fps_cases().items():
        out["fps/" + name] = ref.furthest_point_sampling(T(xyz), m).cpu().numpy()
    for name, (new_xyz, xyz, r, ns) in gc.ball_cases().items():
        out["ball/" + name] = ref.ball_query(T(new_xyz), T(xyz), r, ns).cpu().numpy()
    for name, (u, k) in gc.nn_cases().items():
        d2, idx = ref.three_nn(T(u), T(k))
        out["nn_d2/" + name] = d2.cpu().numpy()
        out["nn_idx/" + name] = idx.cpu().numpy()
    f = gc.feature_case()
    out["feat/group"] = ref.group_points(T(f["feats"]), T(f["idx"])).cpu().numpy()
    out["feat/group_grad"] = ref.group_points_grad(T(f["grad4"]), T(f["idx"]), f["feats"].shape[2]).cpu().numpy()
    out["feat/gather"] = ref.gather_points(T(f["feats"]), T(f["idx1"])).cpu().numpy()
    out["feat/gather_grad"] = ref.gather_points_grad(T(f["grad3"]), T(f["idx1"]), f["feats"].shape[2]).cpu().numpy()
    out["feat/interp"] = ref.three_interpolate(T(f["known"]), T(f["idx3"]), T(f["w3"])).cpu().numpy()
    out["feat/interp_grad"] = ref.three_interpolate_grad(T(f["gradn"]), T(f["idx3"]), T(f["w3"]),
                                                         f["known"].shape[2]).cpu().numpy()
    dst = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(dst, exist_ok=True)
    np.savez_compressed(os.path.join(dst, "native_ops.npz"), **out)
    print("wrote", os.path.join(dst, "native_ops.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
