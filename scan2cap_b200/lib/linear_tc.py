"""nn.Linear on tall inputs (tens of thousands of rows: the EdgeConv message MLP of models/graph_module.py:22-115)
with the backward pass's weight gradient on the tcgen05 kernel of libs2c.so.

Forward and input gradient stay library GEMMs (wide outputs, many tiles: they already fill the GPU).  The weight
gradient dW (out x in) = dY^T X has only a handful of output tiles but a reduction over all rows; the library's fp32
path runs it as a 48-CTA split-K SIMT kernel (~87 us for 20 480 x 256 -> 128), the tensor-core kernel spreads the
rows over every SM (3xTF32, fp32-accurate).  The bias gradient is one column-sum launch."""
import torch
import torch.nn.functional as F
from torch.autograd import Function

from .pointnet2 import _ext_mlp

MIN_ROWS = 2048


def _usable(x, weight):
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[0] >= MIN_ROWS
            and weight.shape[0] <= 256 and weight.shape[0] % 4 == 0 and weight.shape[1] % 4 == 0)


class _LinearTC(Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return F.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        xc = x if (x.stride(1) == 1 and x.stride(0) % 4 == 0) else x.contiguous()
        dx = dy @ weight if ctx.needs_input_grad[0] else None
        dw = _ext_mlp.mlp_layer_bwd_weight_blocked(dy, xc, weight.shape[1]) if ctx.needs_input_grad[1] else None
        db = _ext_mlp.col_sum(dy) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dx, dw, db


def linear(x, weight, bias=None):
    """F.linear(x, weight, bias) with the tensor-core weight gradient when x is a tall CUDA fp32 matrix."""
    if _usable(x, weight) and torch.is_grad_enabled() and (weight.requires_grad or x.requires_grad):
        return _LinearTC.apply(x, weight, bias)
    return F.linear(x, weight, bias)
