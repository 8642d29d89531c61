"""Drop-in for the reference's pybind module ``pointnet2._ext``
(lib/pointnet2/_ext_src/src/bindings.cpp:6-19): same nine function names, argument orders, dtypes,
shapes and error behaviour, implemented by libs2c.so (sm_100a kernels) through its C ABI.

Like the reference, every function requires contiguous CUDA tensors (float32 / int32), allocates its
outputs, and enqueues on the current stream of the tensors' device; violations raise RuntimeError
(the reference raises through AT_ASSERT: include/utils.h:5-25; CPU tensors: "CPU not supported").
"""
import torch

import os

from ..._lib import LIB, call

# ball queries over at least this many points go through the uniform-grid pre-filter; below it the brute-force scan
# is as fast and needs no workspace (both are libs2c kernels with bit-identical results)
GRID_MIN_POINTS = 4096


def _chk(t, name, dtype):
    if not isinstance(t, torch.Tensor):
        raise RuntimeError("%s must be a tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s: CPU not supported" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be a contiguous tensor" % name)
    if t.dtype != dtype:
        raise RuntimeError("%s must be a%s tensor" % (name, " float" if dtype == torch.float32 else "n int"))


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


class _guard(object):
    """Device guard (the reference has none; needed with one process per GPU / multi-device use)."""

    def __init__(self, t):
        self.dev = t.device

    def __enter__(self):
        self.g = torch.cuda.device(self.dev)
        self.g.__enter__()

    def __exit__(self, *a):
        return self.g.__exit__(*a)


def furthest_point_sampling(points, nsamples):
    _chk(points, "points", torch.float32)
    B, N, _ = points.shape
    out = torch.empty((B, nsamples), dtype=torch.int32, device=points.device)
    with _guard(points):
        call("s2c_furthest_point_sampling", points.data_ptr(), B, N, int(nsamples), out.data_ptr(), None,
             _stream(points))
    return out


def furthest_point_sampling_with_xyz(points, nsamples):
    """FPS that also returns the sampled coordinates (B,m,3) (fused gather; not in the reference _ext)."""
    _chk(points, "points", torch.float32)
    B, N, _ = points.shape
    out = torch.empty((B, nsamples), dtype=torch.int32, device=points.device)
    new_xyz = torch.empty((B, nsamples, 3), dtype=torch.float32, device=points.device)
    with _guard(points):
        call("s2c_furthest_point_sampling", points.data_ptr(), B, N, int(nsamples), out.data_ptr(),
             new_xyz.data_ptr(), _stream(points))
    return out, new_xyz


def furthest_point_sampling_chain(points, counts, stream):
    """FPS levels chained on `stream` (a torch.cuda.Stream): level i samples counts[i] points from level i-1's output.
    Outputs are allocated here (on the CURRENT stream's pool) and the kernels are enqueued on `stream`; the caller
    orders the streams (stream.wait_stream(current) before, current.wait_stream(stream) before the first use).
    Lets the coordinate-only sampling of the deeper set-abstraction levels run beside the first level's grouping and
    MLP (backbone_module.py:74-128 evaluates them strictly in sequence)."""
    _chk(points, "points", torch.float32)
    out = []
    cur = points
    with _guard(points):
        for m in counts:
            B, N, _ = cur.shape
            idx = torch.empty((B, m), dtype=torch.int32, device=points.device)
            new_xyz = torch.empty((B, m, 3), dtype=torch.float32, device=points.device)
            call("s2c_furthest_point_sampling", cur.data_ptr(), B, N, int(m), idx.data_ptr(), new_xyz.data_ptr(),
                 stream.cuda_stream)
            out.append((idx, new_xyz))
            cur = new_xyz
    return out


def gather_points(points, idx):
    _chk(points, "points", torch.float32)
    _chk(idx, "idx", torch.int32)
    B, C, N = points.shape
    m = idx.shape[1]
    out = torch.empty((B, C, m), dtype=torch.float32, device=points.device)
    with _guard(points):
        call("s2c_gather_points", points.data_ptr(), idx.data_ptr(), B, C, N, m, out.data_ptr(), _stream(points))
    return out


def gather_points_grad(grad_out, idx, n):
    _chk(grad_out, "grad_out", torch.float32)
    _chk(idx, "idx", torch.int32)
    B, C, m = grad_out.shape
    out = torch.empty((B, C, int(n)), dtype=torch.float32, device=grad_out.device)
    with _guard(grad_out):
        call("s2c_gather_points_grad", grad_out.data_ptr(), idx.data_ptr(), B, C, int(n), m, out.data_ptr(),
             _stream(grad_out))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    _chk(new_xyz, "new_xyz", torch.float32)
    _chk(xyz, "xyz", torch.float32)
    B, M, _ = new_xyz.shape
    n = xyz.shape[1]
    idx = torch.empty((B, M, int(nsample)), dtype=torch.int32, device=new_xyz.device)
    with _guard(new_xyz):
        if n >= GRID_MIN_POINTS and radius > 0:
            ws_bytes = LIB.s2c_ball_query_grid_workspace_bytes(B, n)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xyz.device)
            call("s2c_query_and_group_grid", xyz.data_ptr(), new_xyz.data_ptr(), None, B, n, M, 0, 0, 0, float(radius),
                 int(nsample), 0, 0, idx.data_ptr(), None, ws.data_ptr(), ws_bytes, _stream(new_xyz))
        else:
            call("s2c_ball_query", new_xyz.data_ptr(), xyz.data_ptr(), B, n, M, float(radius), int(nsample),
                 idx.data_ptr(), None, _stream(new_xyz))
    return idx


def group_points(points, idx):
    _chk(points, "points", torch.float32)
    _chk(idx, "idx", torch.int32)
    B, C, N = points.shape
    _, npoints, nsample = idx.shape
    out = torch.empty((B, C, npoints, nsample), dtype=torch.float32, device=points.device)
    with _guard(points):
        call("s2c_group_points", points.data_ptr(), idx.data_ptr(), B, C, N, npoints, nsample, out.data_ptr(),
             _stream(points))
    return out


def group_points_grad(grad_out, idx, n):
    _chk(grad_out, "grad_out", torch.float32)
    _chk(idx, "idx", torch.int32)
    B, C, npoints, nsample = grad_out.shape
    out = torch.empty((B, C, int(n)), dtype=torch.float32, device=grad_out.device)
    with _guard(grad_out):
        call("s2c_group_points_grad", grad_out.data_ptr(), idx.data_ptr(), B, C, int(n), npoints, nsample,
             out.data_ptr(), _stream(grad_out))
    return out


def three_nn(unknowns, knows):
    _chk(unknowns, "unknowns", torch.float32)
    _chk(knows, "knows", torch.float32)
    B, n, _ = unknowns.shape
    m = knows.shape[1]
    dist2 = torch.empty((B, n, 3), dtype=torch.float32, device=unknowns.device)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknowns.device)
    with _guard(unknowns):
        call("s2c_three_nn", unknowns.data_ptr(), knows.data_ptr(), B, n, m, dist2.data_ptr(), idx.data_ptr(),
             _stream(unknowns))
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    _chk(points, "points", torch.float32)
    _chk(idx, "idx", torch.int32)
    _chk(weight, "weight", torch.float32)
    B, C, m = points.shape
    n = idx.shape[1]
    out = torch.empty((B, C, n), dtype=torch.float32, device=points.device)
    with _guard(points):
        call("s2c_three_interpolate", points.data_ptr(), idx.data_ptr(), weight.data_ptr(), B, C, m, n,
             out.data_ptr(), _stream(points))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    _chk(grad_out, "grad_out", torch.float32)
    _chk(idx, "idx", torch.int32)
    _chk(weight, "weight", torch.float32)
    B, C, n = grad_out.shape
    out = torch.empty((B, C, int(m)), dtype=torch.float32, device=grad_out.device)
    with _guard(grad_out):
        call("s2c_three_interpolate_grad", grad_out.data_ptr(), idx.data_ptr(), weight.data_ptr(), B, C, n, int(m),
             out.data_ptr(), _stream(grad_out))
    return out


def ball_query_grid_build(xyz, radius, out=None):
    """The uniform grid of s2c_query_and_group_grid for (xyz (B,n,3), radius), built on the current stream into a uint8
    workspace tensor (returned; `out` reuses a buffer of the same size).  Pass it to query_and_group(..., grid=ws)."""
    _chk(xyz, "xyz", torch.float32)
    B, n, _ = xyz.shape
    ws_bytes = int(LIB.s2c_ball_query_grid_workspace_bytes(B, n))
    ws = out if out is not None else torch.empty(ws_bytes, dtype=torch.uint8, device=xyz.device)
    assert ws.numel() >= ws_bytes and ws.dtype == torch.uint8 and ws.is_contiguous() and ws.data_ptr() % 256 == 0
    with _guard(xyz):
        call("s2c_ball_query_grid_build", xyz.data_ptr(), B, n, float(radius), ws.data_ptr(), ws.numel(), _stream(xyz))
    return ws


def query_and_group(xyz, new_xyz, features, radius, nsample, normalize_xyz, feat_point_major=False,
                    channels_last=False, pad4=False, grid=None):
    """Fused QueryAndGroup.forward (use_xyz=True).  Returns (grouped, idx).
    grid: workspace from ball_query_grid_build(xyz, radius) (the grid was built ahead of time; only the query runs).

    features: None, (B,C,n) [default] or, with feat_point_major, a (B,n,C) view whose last dim is
    contiguous (row stride may exceed C).  grouped: (B,3+C,M,ns) contiguous, or with channels_last the
    same logical shape in torch.channels_last memory format (physically (B,M,ns,3+C)).  With pad4 the rows are laid
    out [x, y, z, 0 | C features | zero pad] with Cp = 4 + ceil4(C) floats: rows AND their feature block are 16-byte
    aligned (TMA operand of the fused MLP; aligned gradient block for the tensor-core dgrad / scatter-add); the
    returned tensor then has Cp channels in that order.
    """
    _chk(xyz, "xyz", torch.float32)
    _chk(new_xyz, "new_xyz", torch.float32)
    B, n, _ = xyz.shape
    M = new_xyz.shape[1]
    C, fptr, fstride, flayout = 0, None, 0, 0
    if features is not None:
        if feat_point_major:
            if not (features.is_cuda and features.dtype == torch.float32 and features.stride(2) == 1
                    and features.stride(0) == n * features.stride(1)):
                raise RuntimeError("features: need a CUDA float (B,n,C) view with unit channel stride")
            C, fstride, flayout = features.shape[2], features.stride(1), 1
        else:
            _chk(features, "features", torch.float32)
            C = features.shape[1]
        fptr = features.data_ptr()
    idx = torch.empty((B, M, int(nsample)), dtype=torch.int32, device=xyz.device)
    pad4 = bool(pad4 and channels_last)
    if channels_last:
        Cp = 4 + (C + 3) // 4 * 4 if pad4 else 3 + C
        grouped = torch.empty((B, M, int(nsample), Cp), dtype=torch.float32, device=xyz.device)
    else:
        grouped = torch.empty((B, 3 + C, M, int(nsample)), dtype=torch.float32, device=xyz.device)
    layout = (2 if pad4 else 1) if channels_last else 0
    with _guard(xyz):
        if grid is not None:
            assert grid.dtype == torch.uint8 and grid.data_ptr() % 256 == 0
            call("s2c_query_and_group_grid_prebuilt", xyz.data_ptr(), new_xyz.data_ptr(), fptr, B, n, M, C, flayout,
                 fstride, float(radius), int(nsample), 1 if normalize_xyz else 0, layout, idx.data_ptr(),
                 grouped.data_ptr(), grid.data_ptr(), grid.numel(), _stream(xyz))
        elif n >= GRID_MIN_POINTS and radius > 0:
            ws_bytes = LIB.s2c_ball_query_grid_workspace_bytes(B, n)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xyz.device)
            call("s2c_query_and_group_grid", xyz.data_ptr(), new_xyz.data_ptr(), fptr, B, n, M, C, flayout, fstride,
                 float(radius), int(nsample), 1 if normalize_xyz else 0, layout, idx.data_ptr(), grouped.data_ptr(),
                 ws.data_ptr(), ws_bytes, _stream(xyz))
        else:
            call("s2c_query_and_group", xyz.data_ptr(), new_xyz.data_ptr(), fptr, B, n, M, C, flayout, fstride,
                 float(radius), int(nsample), 1 if normalize_xyz else 0, layout, idx.data_ptr(), grouped.data_ptr(),
                 _stream(xyz))
    if channels_last:
        grouped = grouped.permute(0, 3, 1, 2)  # logical (B,Cp,M,ns) over channels-last storage
    return grouped, idx
