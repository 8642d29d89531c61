"""Seeded synthetic inputs with the shapes/dtypes of the reference's ``data_dict`` (lib/dataset.py:503-538).

No dataset is shipped with the reference (ScanNet / ScanRefer are licensed), so benchmarks and tests
use a procedural "room": N points sampled uniformly on the surfaces of an 8 m x 6 m x 3 m shell plus
24 axis-aligned boxes standing on the floor.  That gives ScanNet-like *surface* density (about 60 points
inside an r = 0.2 m ball at N = 40 000, i.e. right at SA1's nsample = 64 boundary), and 2 % of the points
are exact duplicates, as produced by the reference's random_sampling-with-replacement
(utils/pc_utils.py:32-40) -- duplicates create exact distance ties and exercise the FPS tie-break rule.
"""
import numpy as np

ROOM = (8.0, 6.0, 3.0)
NUM_BOXES = 24


def _sample_box_surface(rng, lo, hi, n, with_bottom=True):
    """n points uniformly on the surface of the axis-aligned box [lo,hi] -> (xyz (n,3), normals (n,3))."""
    lo = np.asarray(lo, np.float64)
    hi = np.asarray(hi, np.float64)
    d = hi - lo
    # faces: (axis, side)
    faces = [(a, s) for a in range(3) for s in (0, 1)]
    if not with_bottom:
        faces.remove((2, 0))
    areas = np.array([d[(a + 1) % 3] * d[(a + 2) % 3] for a, _ in faces])
    f = rng.choice(len(faces), size=n, p=areas / areas.sum())
    u = rng.random((n, 3))
    pts = lo + u * d
    nrm = np.zeros((n, 3))
    for i, (a, s) in enumerate(faces):
        m = f == i
        pts[m, a] = hi[a] if s else lo[a]
        nrm[m, a] = 1.0 if s else -1.0
    return pts, nrm


def make_boxes(rng, num=NUM_BOXES):
    """(num, 6) boxes as (cx, cy, cz, dx, dy, dz), standing on the floor z=0 inside the room."""
    size = rng.uniform(0.3, 2.0, size=(num, 3))
    size[:, 2] = np.minimum(size[:, 2], 2.5)
    cx = rng.uniform(size[:, 0] / 2 + 0.1, ROOM[0] - size[:, 0] / 2 - 0.1)
    cy = rng.uniform(size[:, 1] / 2 + 0.1, ROOM[1] - size[:, 1] / 2 - 0.1)
    cz = size[:, 2] / 2
    return np.stack([cx, cy, cz, size[:, 0], size[:, 1], size[:, 2]], 1)


def make_scene(n_points, seed=42, dup_frac=0.02, centre=True):
    """One scene: xyz (N,3) f32, normals (N,3) f32, boxes (24,6) f64 (in the same, centred frame)."""
    rng = np.random.default_rng(seed)
    boxes = make_boxes(rng)
    surf = [2 * (ROOM[0] * ROOM[1] + ROOM[0] * ROOM[2] + ROOM[1] * ROOM[2])]
    for b in boxes:
        surf.append(2 * (b[3] * b[5] + b[4] * b[5]) + b[3] * b[4])
    surf = np.array(surf)
    counts = rng.multinomial(n_points, surf / surf.sum())
    pts, nrm = [], []
    p, q = _sample_box_surface(rng, (0, 0, 0), ROOM, counts[0])
    pts.append(p)
    nrm.append(-q)  # room normals point inwards
    for b, c in zip(boxes, counts[1:]):
        p, q = _sample_box_surface(rng, b[:3] - b[3:] / 2, b[:3] + b[3:] / 2, c, with_bottom=False)
        pts.append(p)
        nrm.append(q)
    pts = np.concatenate(pts).astype(np.float32)
    nrm = np.concatenate(nrm).astype(np.float32)
    perm = rng.permutation(n_points)
    pts, nrm = pts[perm], nrm[perm]
    ndup = int(round(dup_frac * n_points))
    if ndup > 0:
        dst = rng.choice(n_points, ndup, replace=False)
        src = rng.integers(0, n_points, ndup)
        pts[dst] = pts[src]
        nrm[dst] = nrm[src]
    if centre:  # ScanNet scans are roughly centred in x/y; keeps |p|^2 > 1e-3 for almost every point
        off = np.array([ROOM[0] / 2, ROOM[1] / 2, 0.0], np.float32)
        pts = pts - off
        boxes = boxes.copy()
        boxes[:, :3] -= off
    return pts, nrm, boxes


def make_point_clouds(batch, n_points, use_normal=False, use_multiview=False, use_height=True, seed=42):
    """point_clouds (B,N,3+C) f32 with the channel order of lib/dataset.py:338-362
    (xyz, [normal], [multiview 128], [height]); also returns the per-scene boxes (B,24,6)."""
    pcs, boxes = [], []
    for b in range(batch):
        xyz, nrm, bx = make_scene(n_points, seed=seed + 1000 * b)
        cols = [xyz]
        if use_normal:
            cols.append(nrm)
        if use_multiview:
            rng = np.random.default_rng(seed + 1000 * b + 7)
            cols.append(np.clip(rng.standard_normal((n_points, 128)).astype(np.float32) * 0.5, -3, 3))
        if use_height:
            floor = np.percentile(xyz[:, 2], 0.99)  # lib/dataset.py:359-362
            cols.append((xyz[:, 2] - floor)[:, None].astype(np.float32))
        pcs.append(np.concatenate(cols, 1))
        boxes.append(bx)
    return np.stack(pcs).astype(np.float32), np.stack(boxes)


def uniform_cube(batch, n_points, seed=42):
    """Uniform points in the unit cube (the microbench sweep's second distribution)."""
    rng = np.random.default_rng(seed)
    return rng.random((batch, n_points, 3), dtype=np.float32)
