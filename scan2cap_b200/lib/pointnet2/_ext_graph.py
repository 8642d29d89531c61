"""Torch-facing wrappers of the graph / caption entry points of libs2c.so (C ABI: include/s2c.h)."""
import torch

from ..._lib import call
from ._ext import _guard, _stream


def knn_adjacency(corners, mask, targets, num_locals, corner_mode, include_self, iou_threshold):
    """corners (B,K,8,3) f64, mask (B,K) int64, targets None or (B,T) int64 ->
    adjacent (B,T,K) f32 [T=K when targets is None], neighbours (B,T,num_locals) int32 (ascending ids)."""
    if not (corners.is_cuda and corners.dtype == torch.float64):
        raise RuntimeError("bbox_corner must be a CUDA float64 tensor")
    corners = corners.contiguous()
    mask = mask.to(torch.int64).contiguous()
    B, K = mask.shape
    if targets is None:
        T, tptr = K, None
    else:
        targets = targets.to(torch.int64).contiguous()
        T, tptr = targets.shape[1], targets.data_ptr()
    adj = torch.empty((B, T, K), dtype=torch.float32, device=corners.device)
    nbr = torch.empty((B, T, int(num_locals)), dtype=torch.int32, device=corners.device)
    with _guard(corners):
        call("s2c_knn_adjacency", corners.data_ptr(), mask.data_ptr(), tptr, B, K, T, int(num_locals),
             1 if corner_mode else 0, 1 if include_self else 0, float(iou_threshold), adj.data_ptr(),
             nbr.data_ptr(), _stream(corners))
    return adj, nbr
