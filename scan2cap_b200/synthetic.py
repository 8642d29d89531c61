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


def make_scene(n_points, seed=42, dup_frac=0.02, centre=True, return_owner=False):
    """One scene: xyz (N,3) f32, normals (N,3) f32, boxes (24,6) f64 (in the same, centred frame)
    [, owner (N) int64: index of the box a point lies on, -1 for the room shell]."""
    rng = np.random.default_rng(seed)
    boxes = make_boxes(rng)
    surf = [2 * (ROOM[0] * ROOM[1] + ROOM[0] * ROOM[2] + ROOM[1] * ROOM[2])]
    for b in boxes:
        surf.append(2 * (b[3] * b[5] + b[4] * b[5]) + b[3] * b[4])
    surf = np.array(surf)
    counts = rng.multinomial(n_points, surf / surf.sum())
    pts, nrm, own = [], [], []
    p, q = _sample_box_surface(rng, (0, 0, 0), ROOM, counts[0])
    pts.append(p)
    nrm.append(-q)  # room normals point inwards
    own.append(np.full(counts[0], -1, np.int64))
    for bi, (b, c) in enumerate(zip(boxes, counts[1:])):
        p, q = _sample_box_surface(rng, b[:3] - b[3:] / 2, b[:3] + b[3:] / 2, c, with_bottom=False)
        pts.append(p)
        nrm.append(q)
        own.append(np.full(c, bi, np.int64))
    pts = np.concatenate(pts).astype(np.float32)
    nrm = np.concatenate(nrm).astype(np.float32)
    own = np.concatenate(own)
    perm = rng.permutation(n_points)
    pts, nrm, own = pts[perm], nrm[perm], own[perm]
    ndup = int(round(dup_frac * n_points))
    if ndup > 0:
        dst = rng.choice(n_points, ndup, replace=False)
        src = rng.integers(0, n_points, ndup)
        pts[dst] = pts[src]
        nrm[dst] = nrm[src]
        own[dst] = own[src]
    if centre:  # ScanNet scans are roughly centred in x/y; keeps |p|^2 > 1e-3 for almost every point
        off = np.array([ROOM[0] / 2, ROOM[1] / 2, 0.0], np.float32)
        pts = pts - off
        boxes = boxes.copy()
        boxes[:, :3] -= off
    if return_owner:
        return pts, nrm, boxes, own
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


# ----------------------------------------------------------------------------------------------------------
# Full training ``data_dict`` (keys / dtypes of lib/dataset.py:503-538 after default collate)
# ----------------------------------------------------------------------------------------------------------
MAX_NUM_OBJ = 128
NUM_GT_BOXES = 12
MAX_DES_LEN = 30  # CONF.TRAIN.MAX_DES_LEN; lang tensors hold MAX_DES_LEN + 2 tokens


def make_vocabulary(num_vocabs=3500, emb_size=300, seed=42):
    """Synthetic stand-in for ScanRefer_vocabulary.json + glove.p (neither is shipped with the reference):
    vocabulary = {"word2idx", "idx2word"} (idx2word keyed by str(i), as the reference expects,
    caption_module.py:561) and embeddings = {word: (emb_size,) float32}.  Index 0 is the pad token."""
    rng = np.random.default_rng(seed)
    words = ["pad_", "unk", "sos", "eos"] + ["w%d" % i for i in range(4, num_vocabs)]
    table = rng.standard_normal((num_vocabs, emb_size)).astype(np.float32)
    vocabulary = {"word2idx": {w: i for i, w in enumerate(words)}, "idx2word": {str(i): w for i, w in enumerate(words)}}
    embeddings = {w: table[i] for i, w in enumerate(words)}
    return vocabulary, embeddings, table


def box_corners(center, size):
    """(…,3),(…,3) float64 -> (…,8,3) float64, corner order of utils/box_util.get_3d_box_batch."""
    sx = np.array([1, 1, -1, -1, 1, 1, -1, -1], np.float64)
    sy = np.array([1, -1, -1, 1, 1, -1, -1, 1], np.float64)
    sz = np.array([1, 1, 1, 1, -1, -1, -1, -1], np.float64)
    sign = np.stack([sx, sy, sz], -1)
    return (size[..., None, :] / 2) * sign + center[..., None, :]


def make_data_dict(batch, n_points, use_normal=False, use_multiview=False, use_height=True, num_vocabs=3500,
                   seed=42, mean_size_arr=None, lang_len=None):
    """Seeded CPU ``data_dict`` (numpy arrays) for a full CapNet forward + loss + backward."""
    if mean_size_arr is None:  # (the reference arm of bench.py loads this file by path and passes its own table)
        from scan2cap_b200.data.scannet.model_util_scannet import MEAN_SIZE_ARR as mean_size_arr
    _, _, table = make_vocabulary(num_vocabs, seed=seed)
    T = MAX_DES_LEN + 2
    d = {k: [] for k in ("point_clouds", "center_label", "size_class_label", "size_residual_label", "sem_cls_label",
                         "box_label_mask", "vote_label", "vote_label_mask", "scene_object_rotations",
                         "scene_object_rotation_masks", "ref_box_corner_label", "gt_box_corner_label",
                         "lang_ids", "lang_len", "lang_feat", "num_bbox")}
    for b in range(batch):
        s = seed + 1000 * b
        rng = np.random.default_rng(s + 3)
        xyz, nrm, boxes, own = make_scene(n_points, seed=s, return_owner=True)
        cols = [xyz]
        if use_normal:
            cols.append(nrm)
        if use_multiview:
            cols.append(np.clip(np.random.default_rng(s + 7).standard_normal((n_points, 128)).astype(np.float32) * 0.5, -3, 3))
        if use_height:
            cols.append((xyz[:, 2] - np.percentile(xyz[:, 2], 0.99))[:, None].astype(np.float32))
        d["point_clouds"].append(np.concatenate(cols, 1).astype(np.float32))
        gt = boxes[:NUM_GT_BOXES]
        center = np.zeros((MAX_NUM_OBJ, 3), np.float32)
        center[:NUM_GT_BOXES] = gt[:, :3]
        cls = np.zeros(MAX_NUM_OBJ, np.int64)
        cls[:NUM_GT_BOXES] = rng.integers(0, 18, NUM_GT_BOXES)
        res = np.zeros((MAX_NUM_OBJ, 3), np.float32)
        res[:NUM_GT_BOXES] = gt[:, 3:] - mean_size_arr[cls[:NUM_GT_BOXES]]
        mask = np.zeros(MAX_NUM_OBJ, np.float32)
        mask[:NUM_GT_BOXES] = 1
        d["center_label"].append(center)
        d["size_class_label"].append(cls)
        d["size_residual_label"].append(res)
        d["sem_cls_label"].append(cls.copy())
        d["box_label_mask"].append(mask)
        d["num_bbox"].append(NUM_GT_BOXES)
        vmask = ((own >= 0) & (own < NUM_GT_BOXES)).astype(np.int64)
        vote = np.zeros((n_points, 3), np.float32)
        vote[vmask == 1] = gt[own[vmask == 1], :3].astype(np.float32) - xyz[vmask == 1]
        d["vote_label"].append(np.tile(vote, (1, 3)))
        d["vote_label_mask"].append(vmask)
        q, _ = np.linalg.qr(rng.standard_normal((MAX_NUM_OBJ, 3, 3)))
        d["scene_object_rotations"].append(q.astype(np.float32))
        d["scene_object_rotation_masks"].append(np.ones(MAX_NUM_OBJ, np.int64))
        corners = np.zeros((MAX_NUM_OBJ, 8, 3), np.float64)
        corners[:NUM_GT_BOXES] = box_corners(gt[:, :3], gt[:, 3:])
        d["gt_box_corner_label"].append(corners)
        d["ref_box_corner_label"].append(corners[b % NUM_GT_BOXES].copy())
        n_tok = int(rng.integers(8, T + 1)) if lang_len is None else int(lang_len)
        ids = np.zeros(T, np.int64)
        ids[:n_tok] = rng.integers(4, num_vocabs, n_tok)
        ids[0], ids[n_tok - 1] = 2, 3
        d["lang_ids"].append(ids)
        d["lang_len"].append(n_tok)
        d["lang_feat"].append(table[ids] * (ids != 0)[:, None])
    out = {k: np.stack(v) if isinstance(v[0], np.ndarray) else np.asarray(v, np.int64) for k, v in d.items()}
    B = batch
    out["heading_class_label"] = np.zeros((B, MAX_NUM_OBJ), np.int64)
    out["heading_residual_label"] = np.zeros((B, MAX_NUM_OBJ), np.float32)
    out["lang_feat"] = out["lang_feat"].astype(np.float32)
    return out


def make_pretrained_data_dict(batch, num_proposals=256, num_valid=64, num_vocabs=3500, seed=42, lang_len=20):
    """Seeded ``data_dict`` of the capnet_pretrained path (BASELINE config 1; keys of lib/dataset_pretrained.py:659-699):
    pre-extracted box features for `num_proposals` boxes of which the first `num_valid` are valid objects.  The
    referred box (ref_box_corner_label) is valid box 3, so good_bbox_masks is not empty."""
    _, _, table = make_vocabulary(num_vocabs, seed=seed)
    T = MAX_DES_LEN + 2
    d = {k: [] for k in ("bbox_feature", "bbox_corner", "bbox_center", "bbox_mask", "bbox_idx", "bbox_corner_label",
                         "bbox_center_label", "ref_box_corner_label", "scene_object_rotations",
                         "scene_object_rotation_masks", "lang_ids", "lang_len", "lang_feat")}
    for b in range(batch):
        rng = np.random.default_rng(seed + 1000 * b + 11)
        center = rng.random((num_proposals, 3)) * np.array(ROOM) - np.array([ROOM[0] / 2, ROOM[1] / 2, 0.0])
        size = rng.uniform(0.3, 2.0, (num_proposals, 3))
        corners = box_corners(center, size)
        mask = np.zeros(num_proposals, np.int64)
        mask[:num_valid] = 1
        d["bbox_feature"].append(rng.standard_normal((num_proposals, 128)).astype(np.float32))
        d["bbox_corner"].append(corners)
        d["bbox_center"].append(center.astype(np.float32))
        d["bbox_mask"].append(mask)
        d["bbox_idx"].append(3)
        d["bbox_corner_label"].append(corners.copy())
        d["bbox_center_label"].append(center.copy())
        d["ref_box_corner_label"].append(corners[3].copy())
        q, _ = np.linalg.qr(rng.standard_normal((num_proposals, 3, 3)))
        d["scene_object_rotations"].append(q.astype(np.float32))
        d["scene_object_rotation_masks"].append(np.ones(num_proposals, np.int64))
        n_tok = int(lang_len)
        ids = np.zeros(T, np.int64)
        ids[:n_tok] = rng.integers(4, num_vocabs, n_tok)
        ids[0], ids[n_tok - 1] = 2, 3
        d["lang_ids"].append(ids)
        d["lang_len"].append(n_tok)
        d["lang_feat"].append((table[ids] * (ids != 0)[:, None]).astype(np.float32))
    return {k: np.stack(v) if isinstance(v[0], np.ndarray) else np.asarray(v, np.int64) for k, v in d.items()}
