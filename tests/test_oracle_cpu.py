"""CPU tests (no GPU): the oracle against (i) the golden vectors produced by the UNMODIFIED reference
CUDA kernels on a B200 (tests/golden/native_ops.npz, made by tests/golden/make_golden_gpu.py),
(ii) the reference's known-answer test, (iii) brute-force numpy restatements."""
import os

import numpy as np
import pytest

import golden_cases as gc

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "native_ops.npz"))


@pytest.mark.parametrize("name", sorted(gc.fps_cases().keys()))
def test_fps_vs_reference_golden(name, oracle):
    xyz, m = gc.fps_cases()[name]
    np.testing.assert_array_equal(oracle.furthest_point_sampling(xyz, m), GOLD["fps/" + name])


@pytest.mark.parametrize("name", sorted(gc.ball_cases().keys()))
def test_ball_query_vs_reference_golden(name, oracle):
    new_xyz, xyz, r, ns = gc.ball_cases()[name]
    np.testing.assert_array_equal(oracle.ball_query(new_xyz, xyz, r, ns), GOLD["ball/" + name])


@pytest.mark.parametrize("name", sorted(gc.nn_cases().keys()))
def test_three_nn_vs_reference_golden(name, oracle):
    u, k = gc.nn_cases()[name]
    d2, idx = oracle.three_nn(u, k)
    np.testing.assert_array_equal(idx, GOLD["nn_idx/" + name])
    np.testing.assert_array_equal(d2, GOLD["nn_d2/" + name])


def test_feature_ops_vs_reference_golden(oracle):
    f = gc.feature_case()
    Nn, m = f["feats"].shape[2], f["known"].shape[2]
    np.testing.assert_array_equal(oracle.group_points(f["feats"], f["idx"]), GOLD["feat/group"])
    np.testing.assert_array_equal(oracle.gather_points(f["feats"], f["idx1"]), GOLD["feat/gather"])
    np.testing.assert_array_equal(oracle.three_interpolate(f["known"], f["idx3"], f["w3"]), GOLD["feat/interp"])
    # gradients are accumulated with atomics in the reference: order-dependent rounding
    np.testing.assert_allclose(oracle.group_points_grad(f["grad4"], f["idx"], Nn), GOLD["feat/group_grad"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(oracle.gather_points_grad(f["grad3"], f["idx1"], Nn), GOLD["feat/gather_grad"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(oracle.three_interpolate_grad(f["gradn"], f["idx3"], f["w3"], m), GOLD["feat/interp_grad"], rtol=1e-5, atol=1e-6)


def test_three_interpolate_reference_kat(oracle):
    """lib/pointnet2/pointnet2_test.py:18-30 (fixed idx / weight)."""
    rng = np.random.default_rng(0)
    feats = rng.standard_normal((1, 2, 4)).astype(np.float32)
    idx = np.array([[[0, 1, 2], [1, 2, 3]]], np.int32)
    w = np.array([[[1, 1, 1], [2, 2, 2]]], np.float32)
    out = oracle.three_interpolate(feats, idx, w)
    f = feats[0]
    want = np.stack([f[:, 0] + f[:, 1] + f[:, 2], 2 * (f[:, 1] + f[:, 2] + f[:, 3])], -1)[None]
    np.testing.assert_allclose(out, want, rtol=1e-5, atol=1e-6)


def test_modules_main_block_case(oracle):
    """pointnet2_modules.py:499-518: randn(2,9,3), radii 5/10 -> every ball holds all 9 points, so the
    neighbour lists are simply the first nsample indices."""
    rng = np.random.default_rng(1)
    xyz = rng.standard_normal((2, 9, 3)).astype(np.float32)
    for r, ns in ((5.0, 6), (10.0, 3)):
        idx = oracle.ball_query(xyz[:, :2].copy(), xyz, r, ns)
        np.testing.assert_array_equal(idx, np.broadcast_to(np.arange(ns, dtype=np.int32), (2, 2, ns)))


def test_fps_properties(oracle):
    rng = np.random.default_rng(3)
    xyz = rng.standard_normal((3, 500, 3)).astype(np.float32) + 2.0
    idx = oracle.furthest_point_sampling(xyz, 100)
    assert (idx[:, 0] == 0).all()
    for b in range(3):
        assert len(set(idx[b].tolist())) == 100  # distinct while distinct points remain
        # each pick maximises the distance to the already-picked set (float64 check, generous tolerance)
        p = xyz[b].astype(np.float64)
        dmin = np.full(500, np.inf)
        for j in range(1, 100):
            dmin = np.minimum(dmin, ((p - p[idx[b, j - 1]]) ** 2).sum(-1))
            assert dmin[idx[b, j]] >= dmin.max() * (1 - 1e-5)


def test_ball_query_bruteforce(oracle):
    rng = np.random.default_rng(4)
    xyz = rng.random((2, 400, 3), dtype=np.float32)
    new_xyz = xyz[:, :50].copy()
    r, ns = 0.2, 12
    idx = oracle.ball_query(new_xyz, xyz, r, ns)
    d2 = ((new_xyz[:, :, None, :].astype(np.float64) - xyz[:, None, :, :]) ** 2).sum(-1)
    for b in range(2):
        for j in range(50):
            hits = np.nonzero(d2[b, j] < r * r - 1e-6)[0][:ns]
            # (points within 1e-6 of the sphere may go either way in float32; none at this seed)
            want = np.full(ns, hits[0] if len(hits) else 0)
            want[:len(hits)] = hits
            np.testing.assert_array_equal(idx[b, j], want)


def test_opt_n_threads(oracle):
    for n, bs in ((1, 1), (2, 2), (3, 2), (31, 16), (32, 32), (511, 256), (512, 512), (40000, 512), (1 << 20, 512)):
        assert oracle.opt_n_threads(n) == bs
