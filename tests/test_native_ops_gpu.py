"""Parity of libs2c's native ops (through the C ABI) against the CPU oracle and, when it is present,
against the UNMODIFIED reference CUDA extension (oracle/_ref).  Integer outputs: bit-exact.
Float outputs of gather/group/interpolate/three_nn: bit-exact too (pure copies / fixed fma order);
atomically accumulated gradients: 1e-5 relative (summation order is unspecified in the reference)."""
import numpy as np
import pytest
import torch

import golden_cases as gc
from scan2cap_b200 import synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------ FPS
@pytest.mark.parametrize("name", sorted(gc.fps_cases().keys()))
def test_fps_golden_cases(name, ext, oracle, ref_ext):
    xyz, m = gc.fps_cases()[name]
    got = N(ext.furthest_point_sampling(T(xyz), m))
    np.testing.assert_array_equal(got, oracle.furthest_point_sampling(xyz, m))
    if ref_ext is not None:
        np.testing.assert_array_equal(got, N(ref_ext.furthest_point_sampling(T(xyz), m)))


@pytest.mark.parametrize("n,m,B", [(40000, 2048, 2), (2048, 1024, 3), (1024, 512, 2), (512, 256, 2), (1024, 256, 2),
                                   (10000, 512, 1), (20000, 300, 1), (8192, 128, 1), (9000, 64, 2), (100000, 64, 1)])
def test_fps_scene_sizes(n, m, B, ext, oracle, ref_ext):
    pc, _ = synthetic.make_point_clouds(B, n, use_height=False, seed=n + m)
    xyz = pc[..., :3].copy()
    idx, new_xyz = ext.furthest_point_sampling_with_xyz(T(xyz), m)
    got = N(idx)
    np.testing.assert_array_equal(got, oracle.furthest_point_sampling(xyz, m))
    np.testing.assert_array_equal(N(new_xyz), np.take_along_axis(xyz, got[..., None].astype(np.int64), 1))
    if ref_ext is not None:
        np.testing.assert_array_equal(got, N(ref_ext.furthest_point_sampling(T(xyz), m)))


@pytest.mark.parametrize("cl", [1, 2, 4, 8, 16])
def test_fps_every_cluster_size(cl, oracle, monkeypatch):
    """Same answer whatever the CTA-cluster decomposition (run in a subprocess: the override is read once)."""
    import subprocess, sys, os
    code = (
        "import numpy as np, torch, sys; sys.path.insert(0, %r)\n"
        "from scan2cap_b200.lib.pointnet2 import _ext\n"
        "from scan2cap_b200 import synthetic\n"
        "pc,_ = synthetic.make_point_clouds(2, 12000, use_height=False, seed=3)\n"
        "idx = _ext.furthest_point_sampling(torch.from_numpy(pc[..., :3].copy()).cuda(), 400)\n"
        "np.save(sys.argv[1], idx.cpu().numpy())\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "idx.npy")
        env = dict(os.environ, S2C_FPS_CLUSTER=str(cl))
        subprocess.check_call([sys.executable, "-c", code, out], env=env)
        got = np.load(out)
    pc, _ = synthetic.make_point_clouds(2, 12000, use_height=False, seed=3)
    np.testing.assert_array_equal(got, oracle.furthest_point_sampling(pc[..., :3].copy(), 400))


# ------------------------------------------------------------------ ball query
@pytest.mark.parametrize("name", sorted(gc.ball_cases().keys()))
def test_ball_query_golden_cases(name, ext, oracle, ref_ext):
    new_xyz, xyz, r, ns = gc.ball_cases()[name]
    got = N(ext.ball_query(T(new_xyz), T(xyz), r, ns))
    np.testing.assert_array_equal(got, oracle.ball_query(new_xyz, xyz, r, ns))
    if ref_ext is not None:
        np.testing.assert_array_equal(got, N(ref_ext.ball_query(T(new_xyz), T(xyz), r, ns)))


@pytest.mark.parametrize("n,M,r,ns", [(40000, 2048, 0.2, 64), (2048, 1024, 0.4, 32), (1024, 512, 0.8, 16),
                                      (512, 256, 1.2, 16), (1024, 256, 0.3, 16), (5000, 33, 0.25, 7)])
def test_ball_query_sa_shapes(n, M, r, ns, ext, oracle, ref_ext):
    pc, _ = synthetic.make_point_clouds(2, n, use_height=False, seed=n)
    xyz = pc[..., :3].copy()
    new_xyz = xyz[:, :M].copy()
    got = N(ext.ball_query(T(new_xyz), T(xyz), r, ns))
    np.testing.assert_array_equal(got, oracle.ball_query(new_xyz, xyz, r, ns))
    if ref_ext is not None:
        np.testing.assert_array_equal(got, N(ref_ext.ball_query(T(new_xyz), T(xyz), r, ns)))


@pytest.mark.parametrize("feat_pm", [False, True])
@pytest.mark.parametrize("cl", [False, True])
@pytest.mark.parametrize("C,normalize", [(0, True), (1, True), (4, False), (37, True), (128, True)])
@pytest.mark.parametrize("n", [3000, 6000])  # below / above S2C_BALL_GRID_MIN: brute-force scan / uniform-grid kernels
def test_query_and_group_fused(feat_pm, cl, C, normalize, n, ext, oracle):
    M, r, ns = 200, 0.3, 32
    pc, _ = synthetic.make_point_clouds(2, n, use_height=False, seed=11)
    xyz = pc[..., :3].copy()
    rng = np.random.default_rng(5)
    new_xyz = xyz[:, rng.permutation(n)[:M]].copy()
    feats_pm = rng.standard_normal((2, n, C)).astype(np.float32)  # point-major
    feats_cm = np.ascontiguousarray(feats_pm.transpose(0, 2, 1))  # (B,C,n)
    # oracle composition = QueryAndGroup.forward (pointnet2_utils.py:334-357)
    idx = oracle.ball_query(new_xyz, xyz, r, ns)
    g_xyz = oracle.group_points(np.ascontiguousarray(xyz.transpose(0, 2, 1)), idx)
    g_xyz = g_xyz - new_xyz.transpose(0, 2, 1)[..., None]
    if normalize:
        g_xyz = g_xyz * (np.float32(1.0) / np.float32(r))
    want = np.concatenate([g_xyz, oracle.group_points(feats_cm, idx)], 1) if C else g_xyz
    f = None
    if C:
        if feat_pm:
            full = T(np.concatenate([xyz, feats_pm], -1))  # like point_clouds (B,n,3+C)
            f = full[..., 3:]
        else:
            f = T(feats_cm)
    grouped, gidx = ext.query_and_group(T(xyz), T(new_xyz), f, r, ns, normalize, feat_point_major=feat_pm,
                                        channels_last=cl)
    np.testing.assert_array_equal(N(gidx), idx)
    assert grouped.shape == (2, 3 + C, M, ns)
    np.testing.assert_array_equal(N(grouped), want)
    if cl:  # padded variant: rows [x, y, z, 0 | features | zero pad], 16-byte aligned rows and feature block
        gp, gidx2 = ext.query_and_group(T(xyz), T(new_xyz), f, r, ns, normalize, feat_point_major=feat_pm,
                                        channels_last=True, pad4=True)
        Cp = 4 + (C + 3) // 4 * 4
        assert gp.shape == (2, Cp, M, ns) and gp.permute(0, 2, 3, 1).is_contiguous()
        np.testing.assert_array_equal(N(gidx2), idx)
        np.testing.assert_array_equal(N(gp[:, :3]), want[:, :3])
        np.testing.assert_array_equal(N(gp[:, 4:4 + C]), want[:, 3:])
        assert float(gp[:, 3].abs().sum()) == 0.0 and float(gp[:, 4 + C:].abs().sum()) == 0.0


# ------------------------------------------------------------------ three_nn / interpolate / gather / group
@pytest.mark.parametrize("name", sorted(gc.nn_cases().keys()))
def test_three_nn_golden_cases(name, ext, oracle, ref_ext):
    u, k = gc.nn_cases()[name]
    d2, idx = ext.three_nn(T(u), T(k))
    od2, oidx = oracle.three_nn(u, k)
    np.testing.assert_array_equal(N(idx), oidx)
    np.testing.assert_array_equal(N(d2), od2)
    if ref_ext is not None:
        rd2, ridx = ref_ext.three_nn(T(u), T(k))
        np.testing.assert_array_equal(N(idx), N(ridx))
        np.testing.assert_array_equal(N(d2), N(rd2))


@pytest.mark.parametrize("n,m", [(512, 256), (1024, 512), (3000, 2500)])
def test_three_nn_fp_shapes(n, m, ext, oracle, ref_ext):
    pc, _ = synthetic.make_point_clouds(2, n + m, use_height=False, seed=n)
    u, k = pc[:, :n, :3].copy(), pc[:, n:, :3].copy()
    d2, idx = ext.three_nn(T(u), T(k))
    od2, oidx = oracle.three_nn(u, k)
    np.testing.assert_array_equal(N(idx), oidx)
    np.testing.assert_array_equal(N(d2), od2)
    if ref_ext is not None:
        rd2, ridx = ref_ext.three_nn(T(u), T(k))
        np.testing.assert_array_equal(N(idx), N(ridx))
        np.testing.assert_array_equal(N(d2), N(rd2))


def test_feature_ops(ext, oracle, ref_ext):
    f = gc.feature_case()
    Nn, m = f["feats"].shape[2], f["known"].shape[2]
    pairs = [
        ("group_points", (f["feats"], f["idx"]), True),
        ("gather_points", (f["feats"], f["idx1"]), True),
        ("three_interpolate", (f["known"], f["idx3"], f["w3"]), True),
        ("group_points_grad", (f["grad4"], f["idx"], Nn), False),
        ("gather_points_grad", (f["grad3"], f["idx1"], Nn), False),
        ("three_interpolate_grad", (f["gradn"], f["idx3"], f["w3"], m), False),
    ]
    for name, args, exact in pairs:
        targs = [T(a) if isinstance(a, np.ndarray) else a for a in args]
        got = N(getattr(ext, name)(*targs))
        want = getattr(oracle, name)(*args)
        if exact:
            np.testing.assert_array_equal(got, want, err_msg=name)
        else:
            np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6, err_msg=name)
        if ref_ext is not None:
            ref = N(getattr(ref_ext, name)(*targs))
            if exact:
                np.testing.assert_array_equal(got, ref, err_msg=name + " vs reference ext")
            else:
                np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6, err_msg=name + " vs reference ext")


def test_three_interpolate_reference_kat(ext):
    """The reference's only test (lib/pointnet2/pointnet2_test.py:18-30): fixed idx/weight; analytic
    values out[:,:,0] = f0+f1+f2, out[:,:,1] = 2*(f1+f2+f3)."""
    torch.manual_seed(0)
    feats = torch.randn(1, 2, 4, device=DEV)
    idx = torch.tensor([[[0, 1, 2], [1, 2, 3]]], dtype=torch.int32, device=DEV)
    w = torch.tensor([[[1., 1., 1.], [2., 2., 2.]]], device=DEV)
    out = ext.three_interpolate(feats, idx, w)
    f = feats[0]
    want = torch.stack([f[:, 0] + f[:, 1] + f[:, 2], 2 * (f[:, 1] + f[:, 2] + f[:, 3])], -1)[None]
    torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-6)


def test_large_group_and_grad_sa1_shape(ext, oracle):
    B, C, n, M, ns = 2, 4, 40000, 2048, 64
    rng = np.random.default_rng(0)
    feats = rng.standard_normal((B, C, n)).astype(np.float32)
    idx = rng.integers(0, n, (B, M, ns)).astype(np.int32)
    np.testing.assert_array_equal(N(ext.group_points(T(feats), T(idx))), oracle.group_points(feats, idx))
    g = rng.standard_normal((B, C, M, ns)).astype(np.float32)
    np.testing.assert_allclose(N(ext.group_points_grad(T(g), T(idx), n)), oracle.group_points_grad(g, idx, n),
                               rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("n,M,r,ns", [(200000, 2048, 0.09, 64), (120000, 1500, 0.12, 16)])
def test_sweep_sizes_fps_and_ball_query(n, M, r, ns, ext, oracle):
    """BASELINE.json configs[4] upper end: N up to 200 000 points per scene (16-CTA clusters, spilled variants)."""
    pc, _ = synthetic.make_point_clouds(1, n, use_height=False, seed=3)
    xyz = pc[..., :3].copy()
    idx, new_xyz = ext.furthest_point_sampling_with_xyz(T(xyz), M)
    np.testing.assert_array_equal(N(idx), oracle.furthest_point_sampling(xyz, M))
    got = ext.ball_query(new_xyz, T(xyz), r, ns)
    np.testing.assert_array_equal(N(got), oracle.ball_query(N(new_xyz), xyz, r, ns))


@pytest.mark.parametrize("n,M,r,ns", [(6000, 300, 5.0, 512), (40000, 512, 0.9, 64), (40000, 256, 5.0, 512),
                                      (5000, 100, 1e-4, 16)])
def test_ball_query_dense_and_empty_balls(n, M, r, ns, ext, oracle, ref_ext):
    """Grid path outside its comfort zone: balls holding thousands of points (more hits than the per-warp hit list:
    the repeated-minimum fallback; MaskVoteNet's r=5 / nsample=512 query, models/mask_votenet.py:145-153) and balls
    holding only the centre itself."""
    pc, _ = synthetic.make_point_clouds(2, n, use_height=False, seed=21)
    xyz = pc[..., :3].copy()
    new_xyz = xyz[:, np.random.default_rng(1).permutation(n)[:M]].copy()
    want = oracle.ball_query(new_xyz, xyz, r, ns)
    np.testing.assert_array_equal(N(ext.ball_query(T(new_xyz), T(xyz), r, ns)), want)
    if ref_ext is not None:
        np.testing.assert_array_equal(want, N(ref_ext.ball_query(T(new_xyz), T(xyz), r, ns)))
    feats = np.random.default_rng(2).standard_normal((2, n, 8)).astype(np.float32)
    grouped, gidx = ext.query_and_group(T(xyz), T(new_xyz), T(feats), r, ns, True, feat_point_major=True,
                                        channels_last=True, pad4=True)
    np.testing.assert_array_equal(N(gidx), want)
    b = np.arange(2)[:, None, None]
    np.testing.assert_array_equal(N(grouped[:, 4:12]).transpose(0, 2, 3, 1), feats[b, want])


@pytest.mark.parametrize("C,c0,ld", [(128, 3, 132), (3, 0, 132), (37, 3, 40), (256, 3, 260)])
def test_group_rows_grad_matches_index_add(C, c0, ld, ext):
    """channels-last scatter-add (gradient of the grouped rows) vs a float64 index_add of the same rows."""
    from scan2cap_b200.lib.pointnet2 import _ext_mlp
    B, n, M, ns = 2, 700, 96, 16
    g = torch.Generator(device="cuda").manual_seed(C)
    rows = torch.randn((B, M * ns, ld), generator=g, device="cuda")
    idx = torch.randint(0, n, (B, M, ns), generator=g, device="cuda", dtype=torch.int32)
    out = _ext_mlp.group_rows_grad(rows, c0, C, idx, n, scale=0.5)
    want = torch.zeros((B, n, C), dtype=torch.float64, device="cuda")
    for b in range(B):
        want[b].index_add_(0, idx[b].reshape(-1).long(), rows[b, :, c0:c0 + C].double() * 0.5)
    # tolerance: fp32 atomics in arbitrary order (the reference's group_points_grad is atomicAdd too)
    torch.testing.assert_close(out.double(), want, rtol=1e-5, atol=1e-5)


def test_errors_raise(ext):
    x = torch.zeros(1, 8, 3)
    with pytest.raises(RuntimeError):
        ext.furthest_point_sampling(x, 4)  # CPU tensor: "CPU not supported"
    with pytest.raises(RuntimeError):
        ext.ball_query(x.cuda(), x.cuda().transpose(1, 2), 0.1, 4)  # non-contiguous
    with pytest.raises(RuntimeError):
        ext.ball_query(x.cuda(), x.cuda(), 0.1, 0)  # nsample out of range -> S2C_ERR_INVALID_ARGUMENT


@pytest.mark.parametrize("ns", [4, 16, 20, 50, 64, 128])
@pytest.mark.parametrize("C,stride_pad", [(132, 0), (132, 4), (36, 0), (252, 0)])
@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
def test_query_and_group_tma_gather_matches_oracle(ns, C, stride_pad, variant, ext, oracle):
    """TMA epilogue of the uniform-grid query (cp.async.bulk.tensor tile::gather4 loads + bulk stores; taken for
    16-byte aligned point-major feature rows with (C+4) % 8 == 0) against the oracle composition AND bit for bit
    against the LDG/STG epilogue, for every ring geometry, tile tails (ns % 8 != 0) and strided source rows."""
    import scan2cap_b200._lib as L
    if variant != 0 and (ns, C, stride_pad) not in ((64, 132, 0), (20, 132, 4), (128, 36, 0)):
        pytest.skip("ring geometries are cross-checked on three shapes")
    B, n, M, r = 2, 6000, 300, 0.3
    pc, _ = synthetic.make_point_clouds(B, n, use_height=False, seed=17)
    xyz = pc[..., :3].copy()
    rng = np.random.default_rng(ns + C)
    new_xyz = xyz[:, rng.permutation(n)[:M]].copy()
    new_xyz[:, -1] += 50.0   # an empty ball: the list stays all-zero, row 0 is gathered nsample times
    feats = rng.standard_normal((B, n, C + stride_pad)).astype(np.float32)
    idx = oracle.ball_query(new_xyz, xyz, r, ns)
    g_xyz = oracle.group_points(np.ascontiguousarray(xyz.transpose(0, 2, 1)), idx)
    g_xyz = (g_xyz - new_xyz.transpose(0, 2, 1)[..., None]) * (np.float32(1.0) / np.float32(r))
    g_f = oracle.group_points(np.ascontiguousarray(feats[..., :C].transpose(0, 2, 1)), idx)
    f = T(feats)[..., :C]   # (B,n,C) view with row stride C + stride_pad
    try:
        L.LIB.s2c_query_and_group_grid_tune(variant)
        gp, gidx = ext.query_and_group(T(xyz), T(new_xyz), f, r, ns, True, feat_point_major=True, channels_last=True,
                                       pad4=True)
        L.LIB.s2c_query_and_group_grid_tune(-1)
        gp_ldg, _ = ext.query_and_group(T(xyz), T(new_xyz), f, r, ns, True, feat_point_major=True, channels_last=True,
                                        pad4=True)
    finally:
        L.LIB.s2c_query_and_group_grid_tune(0)
    np.testing.assert_array_equal(N(gidx), idx)
    assert gp.shape == (B, C + 4, M, ns)
    assert torch.equal(gp, gp_ldg), "TMA and LDG epilogues differ"
    np.testing.assert_array_equal(N(gp[:, :3]), g_xyz)
    np.testing.assert_array_equal(N(gp[:, 4:]), g_f)
    assert float(gp[:, 3].abs().sum()) == 0.0


@pytest.mark.parametrize("length,r", [(3000.0, 0.3), (20000.0, 0.05), (900.0, 0.2)])
def test_ball_query_line_cloud(length, r, ext, oracle):
    """Elongated (line-like) clouds: extent / radius far beyond the 1024 cells per axis the uniform grid allows, so the
    cell edge is coarsened; coordinates up to 2e4 put the fp32 rounding of the cell coordinate near its worst case.
    Neighbour lists must stay bit-exact (the round-1 grid could lose a hit here: VERDICT r01, weak point 5)."""
    rng = np.random.default_rng(int(length))
    n, M, ns = 6000, 500, 16
    xyz = np.zeros((2, n, 3), np.float32)
    xyz[..., 0] = np.sort(rng.uniform(0, length, (2, n)), -1)
    xyz[..., 1:] = rng.uniform(-0.5 * r, 0.5 * r, (2, n, 2))
    # clusters of points at (nearly) one radius from each other: decisions right at the ball's surface
    xyz[:, 1::7, 0] = xyz[:, 0:-1:7, 0][:, :xyz[:, 1::7].shape[1]] + np.float32(r) * np.float32(1 - 1e-6)
    new_xyz = xyz[:, rng.permutation(n)[:M]].copy()
    got = N(ext.ball_query(T(new_xyz), T(xyz), r, ns))
    np.testing.assert_array_equal(got, oracle.ball_query(new_xyz, xyz, r, ns))


@pytest.mark.parametrize("C,ns", [(4, 64), (132, 64), (132, 16)])
def test_prebuilt_grid_equals_one_call(C, ns, ext):
    """s2c_ball_query_grid_build + s2c_query_and_group_grid_prebuilt (the grid built ahead of time, copied between
    buffers, queried twice) == s2c_query_and_group_grid, bit for bit."""
    g = torch.Generator().manual_seed(5)
    B, n, M, r = 3, 20000, 1024, 0.25
    xyz = (torch.rand(B, n, 3, generator=g) * torch.tensor([6.0, 5.0, 2.5])).to(DEV)
    feats = torch.randn(B, n, C, generator=g).to(DEV)
    inds, new_xyz = ext.furthest_point_sampling_with_xyz(xyz, M)
    want_g, want_i = ext.query_and_group(xyz, new_xyz, feats, r, ns, True, feat_point_major=True, channels_last=True, pad4=True)
    ws = ext.ball_query_grid_build(xyz, r)
    ws2 = torch.empty_like(ws)
    ws2.copy_(ws)                      # position independent: what TrainStep does between its staging and static buffers
    for _ in range(2):                 # the work counter is re-armed by every prebuilt call
        got_g, got_i = ext.query_and_group(xyz, new_xyz, feats, r, ns, True, feat_point_major=True, channels_last=True,
                                           pad4=True, grid=ws2)
        assert torch.equal(got_i, want_i)
        assert torch.equal(got_g, want_g)
