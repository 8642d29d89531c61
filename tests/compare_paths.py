"""Comparison implementations used ONLY by the tests (the product has no library-kernel alternative paths)."""
import torch
import torch.nn.functional as F


def bn_rows(x, bn, training):
    """BatchNorm{1d,2d} of a row-major (R, C) matrix with library kernels (statistics over the R rows)."""
    if training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    if bn.momentum is not None:
        mom = bn.momentum
    else:
        mom = 1.0 / float(bn.num_batches_tracked) if training else 0.0
    return F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias,
                        training or not bn.track_running_stats, mom, bn.eps)


def shared_mlp_rows(rows, layers, training):
    """conv1x1(no bias) -> BatchNorm -> ReLU stack on a row-major (R, Cin) matrix: the arithmetic of SharedMLP on a
    (B,C,npoint,nsample) tensor (pytorch_utils.py:11-36, 88-120) with F.linear / F.batch_norm."""
    x = rows
    for conv, bn in layers:
        x = F.linear(x, conv.weight.view(conv.weight.shape[0], -1), conv.bias)
        if bn is not None:
            x = bn_rows(x, bn, training)
        x = F.relu_(x)
    return x
