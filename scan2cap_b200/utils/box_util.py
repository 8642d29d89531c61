"""Device-side mirrors of the two box helpers the hot path uses from the reference's utils/box_util.py:
box3d_iou_batch_tensor (:183-209) and get_3d_box_batch (:360-383, here as a torch function that keeps the
reference's float64 arithmetic but never leaves the GPU)."""
import torch

# corner sign pattern of get_3d_box_batch (box_util.py:374-376): x = +-l/2, y = +-w/2, z = +-h/2
_SX = (1, 1, -1, -1, 1, 1, -1, -1)
_SY = (1, -1, -1, 1, 1, -1, -1, 1)
_SZ = (1, 1, 1, 1, -1, -1, -1, -1)


def get_box3d_min_max_batch_tensor(corner):
    min_coord, _ = corner.min(dim=1)
    max_coord, _ = corner.max(dim=1)
    return (min_coord[:, 0], max_coord[:, 0], min_coord[:, 1], max_coord[:, 1], min_coord[:, 2], max_coord[:, 2])


def box3d_iou_batch_tensor(corners1, corners2):
    """Axis-aligned IoU of (N,8,3) vs (N,8,3) corner sets -> (N)."""
    x_min_1, x_max_1, y_min_1, y_max_1, z_min_1, z_max_1 = get_box3d_min_max_batch_tensor(corners1)
    x_min_2, x_max_2, y_min_2, y_max_2, z_min_2, z_max_2 = get_box3d_min_max_batch_tensor(corners2)
    xA, yA, zA = torch.max(x_min_1, x_min_2), torch.max(y_min_1, y_min_2), torch.max(z_min_1, z_min_2)
    xB, yB, zB = torch.min(x_max_1, x_max_2), torch.min(y_max_1, y_max_2), torch.min(z_max_1, z_max_2)
    inter_vol = (xB - xA).clamp_min(0) * (yB - yA).clamp_min(0) * (zB - zA).clamp_min(0)
    box_vol_1 = (x_max_1 - x_min_1) * (y_max_1 - y_min_1) * (z_max_1 - z_min_1)
    box_vol_2 = (x_max_2 - x_min_2) * (y_max_2 - y_min_2) * (z_max_2 - z_min_2)
    return inter_vol / (box_vol_1 + box_vol_2 - inter_vol + 1e-8)


_SIGN_CACHE = {}


def _corner_signs(dtype, device):
    key = (dtype, str(device))
    if key not in _SIGN_CACHE:
        _SIGN_CACHE[key] = torch.tensor([_SX, _SY, _SZ], dtype=dtype).t().contiguous().to(device)  # (8,3)
    return _SIGN_CACHE[key]


def axis_aligned_corners(box_size, center):
    """get_3d_box_batch for heading 0 (the only heading ScanNet has: model_util_scannet.py:126-136):
    box_size (...,3) f64, center (...,3) f64 -> (...,8,3) f64 = (+-size/2) + center, the same two float64
    operations numpy performs (the rotation by -0.0 rad multiplies by exactly 1 and adds exact zeros)."""
    sign = _corner_signs(box_size.dtype, box_size.device)
    half = box_size / 2
    return half.unsqueeze(-2) * sign + center.unsqueeze(-2)
