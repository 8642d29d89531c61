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


def edgeconv_supported(in_size, out_size):
    return out_size in (64, 128, 256) and (2 * in_size) % 64 == 0 and 2 * in_size <= 512


def _workspace(E, Cin, Cout, backward, device):
    from ..._lib import LIB
    nbytes = int(LIB.s2c_edgeconv_workspace_bytes(int(E), int(Cin), int(Cout), 1 if backward else 0))
    if nbytes < 0:
        raise RuntimeError("s2c_edgeconv_workspace_bytes: bad sizes")
    return torch.empty(nbytes + 256, dtype=torch.uint8, device=device)


def _aligned_ptr(ws):
    return (ws.data_ptr() + 255) // 256 * 256


def edgeconv_fwd(x, row, col, edge_mask, W1, b1, W2, b2, want_agg):
    """One EdgeConv layer (include/s2c.h: s2c_edgeconv_fwd).  x (Nn, Cin) fp32, row/col (E) int64, edge_mask (E) bool or
    None -> (z (E, 2Cin), Y1 (E, Cout), msg (E, Cout) masked messages, agg (Nn, Cout) or None)."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
    Nn, Cin = x.shape
    Cout, E = W1.shape[0], row.shape[0]
    assert row.dtype == torch.int64 and col.dtype == torch.int64 and row.is_contiguous() and col.is_contiguous()
    assert W1.is_contiguous() and W2.is_contiguous() and W1.shape == (Cout, 2 * Cin) and W2.shape == (Cout, Cout)
    z = torch.empty((E, 2 * Cin), dtype=torch.float32, device=x.device)
    Y1 = torch.empty((E, Cout), dtype=torch.float32, device=x.device)
    msg = torch.empty((E, Cout), dtype=torch.float32, device=x.device)
    agg = torch.empty((Nn, Cout), dtype=torch.float32, device=x.device) if want_agg else None
    mask8 = None
    if edge_mask is not None:
        mask8 = edge_mask.contiguous().view(torch.uint8) if edge_mask.dtype == torch.bool else edge_mask.to(torch.uint8)
    ws = _workspace(E, Cin, Cout, False, x.device)
    with _guard(x):
        call("s2c_edgeconv_fwd", x.data_ptr(), Nn, Cin, row.data_ptr(), col.data_ptr(),
             mask8.data_ptr() if mask8 is not None else None, E, W1.data_ptr(), b1.data_ptr(), W2.data_ptr(),
             b2.data_ptr(), Cout, z.data_ptr(), Y1.data_ptr(), msg.data_ptr(), agg.data_ptr() if want_agg else None,
             _aligned_ptr(ws), _stream(x))
    return z, Y1, msg, agg, mask8


def edgeconv_bwd(dagg, dmsg, Nn, row, col, mask8, W1, b1, W2, z, Y1, want_dx):
    """-> (dx (Nn, Cin) or None, dW1, db1, dW2, db2)  (include/s2c.h: s2c_edgeconv_bwd)."""
    E, Cout = Y1.shape
    Cin = z.shape[1] // 2
    dev = z.device
    dx = torch.empty((Nn, Cin), dtype=torch.float32, device=dev) if want_dx else None
    dW1 = torch.empty((Cout, 2 * Cin), dtype=torch.float32, device=dev)
    dW2 = torch.empty((Cout, Cout), dtype=torch.float32, device=dev)
    db = torch.empty((2, Cout), dtype=torch.float32, device=dev)
    ws = _workspace(E, Cin, Cout, True, dev)
    ptr = lambda t: t.data_ptr() if t is not None else None
    with _guard(z):
        call("s2c_edgeconv_bwd", ptr(dagg), ptr(dmsg), Nn, Cin, row.data_ptr(), col.data_ptr(), ptr(mask8), E,
             W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), Cout, z.data_ptr(), Y1.data_ptr(), ptr(dx), dW1.data_ptr(),
             db[0].data_ptr(), dW2.data_ptr(), db[1].data_ptr(), _aligned_ptr(ws), _stream(z))
    return dx, dW1, db[0], dW2, db[1]
