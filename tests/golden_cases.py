"""Seeded inputs of the golden-vector cases for the nine native ops (shared by
tests/golden/make_golden_gpu.py, which runs the UNMODIFIED reference kernels on them on a B200,
and by the tests that replay them against the oracle and against libs2c)."""
import numpy as np

from scan2cap_b200 import synthetic


def fps_cases():
    rng = np.random.default_rng(1234)
    room, _, _ = synthetic.make_scene(4096, seed=5)
    c = {}
    c["room4096_256"] = (room[None].copy(), 256)                       # bs=512, duplicates -> ties
    c["rand700_64"] = (rng.standard_normal((2, 700, 3)).astype(np.float32), 64)   # bs=512, ragged tail
    c["rand300_50"] = (rng.standard_normal((1, 300, 3)).astype(np.float32), 50)   # bs=256
    c["rand37_37"] = (rng.standard_normal((1, 37, 3)).astype(np.float32), 37)     # bs=32, m == N
    near0 = rng.standard_normal((1, 600, 3)).astype(np.float32)
    near0[0, ::3] *= 0.01                                               # |p|^2 <= 1e-3 -> skipped points
    near0[0, 0] = 0.0                                                   # the seed point itself is skipped
    c["skip600_128"] = (near0, 128)
    dup = np.repeat(rng.standard_normal((1, 40, 3)).astype(np.float32), 16, axis=1)  # 640 pts, 40 distinct
    c["dup640_100"] = (dup[:, rng.permutation(640)], 100)               # m > #distinct -> all-tie tail
    c["allzero64_8"] = (np.zeros((1, 64, 3), np.float32), 8)            # no admissible point at all
    grid = np.stack(np.meshgrid(*[np.arange(8, dtype=np.float32)] * 3, indexing="ij"), -1).reshape(1, 512, 3) + 1
    c["lattice512_200"] = (grid, 200)                                   # massive exact ties
    return c


def ball_cases():
    rng = np.random.default_rng(4321)
    room, _, _ = synthetic.make_scene(4096, seed=6)
    xyz = room[None].copy()
    c = {}
    c["room_r02_ns16"] = (xyz[:, :128].copy(), xyz, 0.2, 16)
    c["room_r04_ns64"] = (xyz[:, 100:164].copy(), xyz, 0.4, 64)
    far = xyz[:, :32].copy()
    far[0, ::2] += 100.0                                                # empty balls -> zeros
    c["empty_r02_ns8"] = (far, xyz, 0.2, 8)
    cube = rng.random((2, 1000, 3), dtype=np.float32)
    c["cube_r015_ns32"] = (cube[:, :77].copy(), cube, 0.15, 32)
    c["cube_r3_ns5"] = (cube[:, :9].copy(), cube, 3.0, 5)               # everything inside: idx = 0..4
    return c


def nn_cases():
    rng = np.random.default_rng(99)
    c = {}
    c["u64_k32"] = (rng.standard_normal((2, 64, 3)).astype(np.float32), rng.standard_normal((2, 32, 3)).astype(np.float32))
    c["u10_k2"] = (rng.standard_normal((1, 10, 3)).astype(np.float32), rng.standard_normal((1, 2, 3)).astype(np.float32))
    lat = np.stack(np.meshgrid(*[np.arange(4, dtype=np.float32)] * 3, indexing="ij"), -1).reshape(1, 64, 3)
    c["lattice_ties"] = (lat + 0.5, lat)                                # 8 equidistant neighbours each
    return c


def feature_case():
    rng = np.random.default_rng(7)
    B, C, N, npoint, ns = 2, 5, 200, 16, 8
    feats = rng.standard_normal((B, C, N)).astype(np.float32)
    idx = rng.integers(0, N, (B, npoint, ns)).astype(np.int32)
    idx1 = rng.integers(0, N, (B, npoint)).astype(np.int32)
    grad4 = rng.standard_normal((B, C, npoint, ns)).astype(np.float32)
    grad3 = rng.standard_normal((B, C, npoint)).astype(np.float32)
    m, n = 24, 40
    known = rng.standard_normal((B, C, m)).astype(np.float32)
    idx3 = rng.integers(0, m, (B, n, 3)).astype(np.int32)
    w3 = rng.random((B, n, 3)).astype(np.float32)
    w3 /= w3.sum(-1, keepdims=True)
    gradn = rng.standard_normal((B, C, n)).astype(np.float32)
    return dict(feats=feats, idx=idx, idx1=idx1, grad4=grad4, grad3=grad3, known=known, idx3=idx3, w3=w3,
                gradn=gradn)
