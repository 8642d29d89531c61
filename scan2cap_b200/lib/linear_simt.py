"""nn.Linear on the plain-fp32 GEMM kernels of libs2c.so (s2c_gemm / s2c_gemm_tn), forward and backward.

The caption module's Linear layers (models/caption_module.py:216-240 -- map_feat, the hoisted word / target terms of
map_topdown, classifier) have widths (300, 812, 3500) that are not multiples of the tensor-core kernels' 64-column
tiles and only a few hundred rows; they were the last library (cuBLAS) GEMMs of the training step.  Weights may be
column slices of a wider matrix (row stride > width); inputs of any stride."""
import torch
from torch.autograd import Function

from .._lib import call


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def gemm(A, B_kn_strides, Bt, M, N, K, bias=None, relu=False, out=None):
    """out (M, N) = A (M, K) @ B (K, N) with B(k, n) = Bt.data_ptr()[k * sbk + n * sbn], (sbk, sbn) = B_kn_strides."""
    assert A.is_cuda and A.dtype == torch.float32 and Bt.dtype == torch.float32
    out = torch.empty((M, N), dtype=torch.float32, device=A.device) if out is None else out
    with torch.cuda.device(A.device):
        call("s2c_gemm", A.data_ptr(), A.stride(0), A.stride(1), Bt.data_ptr(), int(B_kn_strides[0]), int(B_kn_strides[1]),
             bias.data_ptr() if bias is not None else None, int(bool(relu)), int(M), int(N), int(K), out.data_ptr(),
             out.stride(0), _stream(A))
    return out


class _Linear(Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        """x (R, K), weight (N, K) [any strides], bias (N) or None -> x @ weight^T + bias."""
        R, K = x.shape
        N = weight.shape[0]
        out = gemm(x, (weight.stride(1), weight.stride(0)), weight, R, N, K, bias=bias)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        R, K = x.shape
        N = weight.shape[0]
        dy = dy if dy.stride(1) == 1 else dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:   # dx = dy @ weight: B(k = n, n = k) = weight[n, k]
            dx = gemm(dy, (weight.stride(0), weight.stride(1)), weight, R, K, N)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            xr = x if x.stride(1) == 1 else x.contiguous()
            dw = torch.empty((N, K), dtype=torch.float32, device=x.device)
            db = torch.empty((N,), dtype=torch.float32, device=x.device) if ctx.has_bias else None
            with torch.cuda.device(x.device):
                call("s2c_gemm_tn", dy.data_ptr(), dy.stride(0), xr.data_ptr(), xr.stride(0), R, N, K, dw.data_ptr(), K,
                     db.data_ptr() if db is not None else None, _stream(x))
        return dx, dw, db


def linear(x, weight, bias=None):
    """F.linear(x, weight, bias) for CUDA fp32 tensors on the libs2c kernels; x (..., K) -> (..., N)."""
    if not (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and weight.dim() == 2
            and weight.shape[1] == x.shape[-1]):
        raise RuntimeError("linear_simt.linear: CUDA fp32 x (..., K) and weight (N, K) expected; got %s %s / %s"
                           % (x.device, tuple(x.shape), tuple(weight.shape)))
    x2 = x.reshape(-1, x.shape[-1])
    y = _Linear.apply(x2, weight, bias)
    return y.view(*x.shape[:-1], weight.shape[0])
