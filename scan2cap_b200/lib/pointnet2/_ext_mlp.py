"""Torch-facing wrappers of the grouped-MLP (tensor-core) entry points of libs2c.so (C ABI: include/s2c.h)."""
import torch

from ..._lib import call
from ._ext import _guard, _stream


# Two libs2c kernels implement a layer: the warp-specialised TMA-fed one (mlp2.cu; output widths 64/128/256, 16-byte
# aligned rows -- every layer of the CapNet stacks) and the single-role one (mlp.cu) that takes the remaining shapes
# (unaligned rows, widths that are other multiples of 16).  `version` is a test hook, not a product switch.
KERNEL_VERSION = 2


def mlp_layer_fwd(A, W, pro_scale=None, pro_shift=None, want_stats=True, K=None, version=None, out=None, col0=0,
                  stats=None):
    """A (R, lda) fp32 (row stride lda >= K), W (N, K) -> C (R, N) = relu(A*scale+shift) @ W^T [no prologue when
    scale is None], plus float64 column sums / sums of squares of C when want_stats.  With `out` (R, >= col0+N) the
    result is written to out[:, col0:col0+N].  stats: a zero-filled float64 (2, >=N) buffer to accumulate the statistics
    into (the caller zero-fills several layers' buffers with one launch); allocated here otherwise."""
    assert A.is_cuda and A.dtype == torch.float32 and A.dim() == 2 and A.stride(1) == 1
    W = W.contiguous()
    N, Kw = W.shape
    K = Kw if K is None else K
    assert K == Kw and A.shape[1] >= K
    R, lda = A.shape[0], A.stride(0)
    C = torch.empty((R, N), dtype=torch.float32, device=A.device) if out is None else out
    assert C.stride(1) == 1 and C.shape[0] == R and C.shape[1] >= col0 + N and col0 % 4 == 0
    ldc, cptr = C.stride(0), C.data_ptr() + 4 * col0
    s1 = s2 = None
    if want_stats:
        if stats is None:
            stats = torch.zeros((2, N), dtype=torch.float64, device=A.device)
        s1, s2 = stats[0, :N], stats[1, :N]
    version = KERNEL_VERSION if version is None else version
    v2_ok = (N in (64, 128, 256) and K % 4 == 0 and lda % 4 == 0 and A.data_ptr() % 16 == 0 and K >= 4)
    with _guard(A):
        if version == 2 and v2_ok:
            wprep = torch.empty(((K + 31) // 32) * N * 256, dtype=torch.uint8, device=A.device)
            call("s2c_mlp_layer_fwd_v2", A.data_ptr(), lda, R, K,
                 pro_scale.data_ptr() if pro_scale is not None else None,
                 pro_shift.data_ptr() if pro_shift is not None else None,
                 W.data_ptr(), N, cptr, ldc,
                 s1.data_ptr() if want_stats else None, s2.data_ptr() if want_stats else None, wprep.data_ptr(),
                 _stream(A))
        else:
            call("s2c_mlp_layer_fwd", A.data_ptr(), lda, R, K,
                 pro_scale.data_ptr() if pro_scale is not None else None,
                 pro_shift.data_ptr() if pro_shift is not None else None,
                 W.data_ptr(), N, cptr, ldc,
                 s1.data_ptr() if want_stats else None, s2.data_ptr() if want_stats else None, _stream(A))
    return (C, s1, s2) if want_stats else C


def pool_fwd(Y, G, ns, scale, shift, want_argmax=True):
    """Y (G*ns, N) -> out (G, N) = max_s relu(Y*scale+shift), argmax (G, N) int32."""
    R, N = Y.shape
    assert R == G * ns and Y.stride(1) == 1
    out = torch.empty((G, N), dtype=torch.float32, device=Y.device)
    am = torch.empty((G, N), dtype=torch.int32, device=Y.device) if want_argmax else None
    with _guard(Y):
        call("s2c_pool_fwd", Y.data_ptr(), Y.stride(0), G, ns, N, scale.data_ptr(), shift.data_ptr(), out.data_ptr(),
             am.data_ptr() if want_argmax else None, _stream(Y))
    return out, am


def pool_bwd_stats(dpool, argmax, Y, ns, scale, shift, stats=None):
    """-> float64 (sum_g [N], sum_gy [N]) of the ReLU-masked pooled gradient routed to the arg-max elements.
    stats: optional zero-filled float64 (2, >=N) accumulator."""
    G, N = dpool.shape
    if stats is None:
        stats = torch.zeros((2, N), dtype=torch.float64, device=Y.device)
    s1, s2 = stats[0, :N], stats[1, :N]
    with _guard(Y):
        call("s2c_pool_bwd_stats", dpool.data_ptr(), argmax.data_ptr(), Y.data_ptr(), Y.stride(0), G, ns, N,
             scale.data_ptr(), shift.data_ptr(), s1.data_ptr(), s2.data_ptr(), _stream(Y))
    return s1, s2


def bwd_data_supported(K, N):
    return KERNEL_VERSION == 2 and N in (64, 128, 256) and K % 4 == 0 and K >= 4


def mlp_layer_bwd_data(Y, a, b, c, W, Yprev, prev_scale, prev_shift, G=None, dpool=None, argmax=None, ns=1,
                       last_scale=None, last_shift=None, want_dY=True, stats=None):
    """One layer's fused backward-data pass (see include/s2c.h).  Y (R,K) pre-BN output of layer l; W (K,N) its
    weight; Yprev (R,N) pre-BN output of layer l-1.  Gradient in: dense G (R,K) or (dpool, argmax) (R/ns, K).
    Returns g_prev (R,N), dY (R,K) or None, sum_g (N) f64, sum_gy (N) f64."""
    R, K = Y.shape
    N = Yprev.shape[1]
    W = W.contiguous()
    assert W.shape == (K, N)
    gprev = torch.empty((R, N), dtype=torch.float32, device=Y.device)
    dY = torch.empty((R, K), dtype=torch.float32, device=Y.device) if want_dY else None
    if stats is None:
        stats = torch.zeros((2, N), dtype=torch.float64, device=Y.device)
    s1, s2 = stats[0, :N], stats[1, :N]
    wprep = torch.empty(((K + 31) // 32) * N * 256, dtype=torch.uint8, device=Y.device)
    ptr = lambda t: t.data_ptr() if t is not None else None
    with _guard(Y):
        call("s2c_mlp_layer_bwd_data", ptr(G), G.stride(0) if G is not None else 0, Y.data_ptr(), Y.stride(0), R, K,
             a.data_ptr(), b.data_ptr(), c.data_ptr(), ptr(dpool), ptr(argmax), int(ns), ptr(last_scale), ptr(last_shift),
             W.data_ptr(), N, Yprev.data_ptr(), Yprev.stride(0), prev_scale.data_ptr(), prev_shift.data_ptr(),
             gprev.data_ptr(), N, ptr(dY), s1.data_ptr(), s2.data_ptr(), wprep.data_ptr(), _stream(Y))
    return gprev, dY, s1, s2


def mlp_layer_bwd_input(G, Y, a, b, c, W, col0, N, out, want_dY=True):
    """First-layer input gradient (see include/s2c.h): out[:, col0:col0+N] = (a*G + b*Y + c) @ W[:, col0:col0+N];
    W (K, ldw) the layer's weight in the column layout of `out` (R, ld).  Returns dY (R, K) or None."""
    R, K = Y.shape
    assert W.shape[0] == K and W.stride(1) == 1 and out.stride(1) == 1 and col0 % 4 == 0 and N in (64, 128, 256)
    dY = torch.empty((R, K), dtype=torch.float32, device=Y.device) if want_dY else None
    wprep = torch.empty(((K + 31) // 32) * N * 256, dtype=torch.uint8, device=Y.device)
    with _guard(Y):
        call("s2c_mlp_layer_bwd_input", G.data_ptr(), G.stride(0), Y.data_ptr(), Y.stride(0), R, K, a.data_ptr(),
             b.data_ptr(), c.data_ptr(), W.data_ptr() + 4 * col0, W.stride(0), int(N), out.data_ptr() + 4 * col0,
             out.stride(0), dY.data_ptr() if want_dY else None, wprep.data_ptr(), _stream(Y))
    return dY


def bwd_weight_supported(C, P, *lds):
    MH, NB = (C + 127) // 128, (P + 31) // 32
    return (KERNEL_VERSION == 2 and C <= 256 and NB <= 9 and MH * NB <= 16 and 4 * MH + NB <= 13
            and all(ld % 4 == 0 for ld in lds))


def mlp_layer_bwd_weight(dY, X, P, xs=None, xh=None, a=None, b=None, c=None, Y=None, out=None, col0=0):
    """dW (C, P) = sum_r dY[r]^T X'[r]; dY (R, C) dense or (with a, b, c, Y) formed as a*dY + b*Y + c; X (R, >=P)
    with optional relu(X*xs+xh) prologue.  With `out` (C, >= col0+P, zero-filled by the caller) the block is written
    to out[:, col0:col0+P] from the columns X[:, col0:col0+P]."""
    R, C = dY.shape
    dW = torch.zeros((C, P), dtype=torch.float32, device=dY.device) if out is None else out
    ptr = lambda t, o=0: t.data_ptr() + 4 * o if t is not None else None
    with _guard(dY):
        call("s2c_mlp_layer_bwd_weight", dY.data_ptr(), dY.stride(0), ptr(Y), Y.stride(0) if Y is not None else 0,
             ptr(a), ptr(b), ptr(c), ptr(X, col0), X.stride(0), ptr(xs, col0), ptr(xh, col0), R, C, int(P),
             ptr(dW, col0), dW.stride(0), _stream(dY))
    return dW


def wgrad_blocked_supported(C, P, *lds):
    return KERNEL_VERSION == 2 and C <= 256 and C % 4 == 0 and P % 4 == 0 and all(ld % 4 == 0 for ld in lds)


def mlp_layer_bwd_weight_blocked(dY, X, P, xs=None, xh=None, out=None):
    """The same weight gradient for any width P: column blocks of the widest shape the tensor-core kernel holds
    (256 columns for C <= 128 output channels, 128 for C <= 256), each one launch over all rows."""
    R, C = dY.shape
    step = 256 if C <= 128 else 128
    dW = torch.zeros((C, P), dtype=torch.float32, device=dY.device) if out is None else out
    for c0 in range(0, P, step):
        mlp_layer_bwd_weight(dY, X, min(step, P - c0), xs, xh, out=dW, col0=c0)
    return dW


def col_sum(A):
    """(R, M) -> (M): column sums (bias gradients)."""
    R, M = A.shape
    assert A.is_cuda and A.dtype == torch.float32 and A.stride(1) == 1
    out = torch.empty((M,), dtype=torch.float32, device=A.device)
    with _guard(A):
        call("s2c_col_sum", A.data_ptr(), A.stride(0), R, M, out.data_ptr(), _stream(A))
    return out


def bn_finalize(s1, s2, R, bn, training):
    """One launch: BatchNorm statistics of a layer -> (mean f64, invstd f64, scale f32, shift f32) [N]; updates the
    module's running statistics in training mode (s1/s2 = float64 column sums; None -> running statistics)."""
    N = bn.weight.shape[0]
    dev = bn.weight.device
    use_batch = bool(training or not bn.track_running_stats)
    update = bool(training and bn.track_running_stats)
    f64 = torch.empty((2, N), dtype=torch.float64, device=dev)
    f32 = torch.empty((2, N), dtype=torch.float32, device=dev)
    ptr = lambda t: t.data_ptr() if t is not None else None
    mom = bn.momentum if bn.momentum is not None else -1.0  # < 0: cumulative average, 1 / num_batches_tracked
    with _guard(bn.weight):
        call("s2c_bn_finalize", ptr(s1) if use_batch else None, ptr(s2) if use_batch else None, int(R), N,
             bn.weight.data_ptr(), bn.bias.data_ptr(), float(bn.eps), float(mom), int(use_batch), int(update),
             ptr(bn.running_mean), ptr(bn.running_var), ptr(bn.num_batches_tracked) if update else None,
             f64[0].data_ptr(), f64[1].data_ptr(), f32[0].data_ptr(), f32[1].data_ptr(), _stream(bn.weight))
    return f64[0], f64[1], f32[0], f32[1]


def bn_backward_coeffs(sum_g, sum_gy, mean, invstd, gamma, R, batch_stats):
    """One launch: -> (grad_gamma, grad_beta, a, b, c) fp32 [N] with dY = a*g + b*y + c."""
    N = gamma.shape[0]
    out = torch.empty((5, N), dtype=torch.float32, device=gamma.device)
    with _guard(gamma):
        call("s2c_bn_backward_coeffs", sum_g.data_ptr(), sum_gy.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
             gamma.data_ptr(), int(R), N, int(bool(batch_stats)), out[0].data_ptr(), out[1].data_ptr(),
             out[2].data_ptr(), out[3].data_ptr(), out[4].data_ptr(), _stream(gamma))
    return out[0], out[1], out[2], out[3], out[4]


def group_rows_grad(rows, c0, C, idx, n, scale=1.0):
    """rows (B, T, ld) channels-last gradient of a grouped tensor, idx (B, ...) with T entries per scene ->
    (B, n, C) point-major gradient of the gathered tensor (scatter-add of channels [c0, c0+C), times scale)."""
    B, T, ld = rows.shape
    assert rows.is_cuda and rows.dtype == torch.float32 and rows.is_contiguous() and idx.dtype == torch.int32
    assert idx.is_contiguous() and idx.numel() == B * T
    out = torch.empty((B, n, C), dtype=torch.float32, device=rows.device)
    with _guard(rows):
        call("s2c_group_rows_grad", rows.data_ptr(), ld, int(c0), int(C), idx.data_ptr(), B, T, int(n), float(scale),
             out.data_ptr(), _stream(rows))
    return out
