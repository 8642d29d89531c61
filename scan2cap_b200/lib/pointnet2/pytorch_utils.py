"""Mirror of the reference's lib/pointnet2/pytorch_utils.py (SharedMLP :11-36, _ConvBase :67-120,
Conv1d/Conv2d :123-188, BNMomentumScheduler :271-298) -- same class names, constructor arguments and
state-dict keys (``layer{i}.conv.weight``, ``layer{i}.bn.bn.{weight,bias,running_mean,running_var,
num_batches_tracked}``), so the reference's checkpoints load unchanged.

Only what CapNet instantiates is provided: 1x1 convolutions WITHOUT bias followed by BatchNorm and ReLU
(bn=True) or with bias (bn=False).  ``SharedMLP.layer_params()`` exposes the per-layer tensors to the fused
grouped-MLP path of pointnet2_modules.py.
"""
import torch
import torch.nn as nn


class _BNBase(nn.Sequential):
    def __init__(self, in_size, batch_norm=None, name=""):
        super().__init__()
        self.add_module(name + "bn", batch_norm(in_size))
        nn.init.constant_(self[0].weight, 1.0)
        nn.init.constant_(self[0].bias, 0)


class BatchNorm1d(_BNBase):
    def __init__(self, in_size, *, name=""):
        super().__init__(in_size, batch_norm=nn.BatchNorm1d, name=name)


class BatchNorm2d(_BNBase):
    def __init__(self, in_size, name=""):
        super().__init__(in_size, batch_norm=nn.BatchNorm2d, name=name)


class _ConvBase(nn.Sequential):
    def __init__(self, in_size, out_size, conv, batch_norm, *, activation, bn, init, bias, preact, name):
        super().__init__()
        bias = bias and (not bn)
        conv_unit = conv(in_size, out_size, kernel_size=1, stride=1, padding=0, bias=bias)
        init(conv_unit.weight)
        if bias:
            nn.init.constant_(conv_unit.bias, 0)
        bn_unit = batch_norm(in_size if preact else out_size) if bn else None
        if preact:
            if bn:
                self.add_module(name + "bn", bn_unit)
            if activation is not None:
                self.add_module(name + "activation", activation)
        self.add_module(name + "conv", conv_unit)
        if not preact:
            if bn:
                self.add_module(name + "bn", bn_unit)
            if activation is not None:
                self.add_module(name + "activation", activation)


class Conv1d(_ConvBase):
    def __init__(self, in_size, out_size, *, kernel_size=1, stride=1, padding=0, activation=nn.ReLU(inplace=True),
                 bn=False, init=nn.init.kaiming_normal_, bias=True, preact=False, name=""):
        assert kernel_size == 1 and stride == 1 and padding == 0, "only pointwise convolutions are supported"
        super().__init__(in_size, out_size, nn.Conv1d, BatchNorm1d, activation=activation, bn=bn, init=init,
                         bias=bias, preact=preact, name=name)


class Conv2d(_ConvBase):
    def __init__(self, in_size, out_size, *, kernel_size=(1, 1), stride=(1, 1), padding=(0, 0),
                 activation=nn.ReLU(inplace=True), bn=False, init=nn.init.kaiming_normal_, bias=True,
                 preact=False, name=""):
        assert tuple(kernel_size) == (1, 1) and tuple(stride) == (1, 1) and tuple(padding) == (0, 0), \
            "only pointwise convolutions are supported"
        super().__init__(in_size, out_size, nn.Conv2d, BatchNorm2d, activation=activation, bn=bn, init=init,
                         bias=bias, preact=preact, name=name)


class SharedMLP(nn.Sequential):
    def __init__(self, args, *, bn=False, activation=nn.ReLU(inplace=True), preact=False, first=False, name=""):
        super().__init__()
        for i in range(len(args) - 1):
            plain = (not first or not preact or (i != 0))
            self.add_module(name + "layer{}".format(i),
                            Conv2d(args[i], args[i + 1], bn=plain and bn, activation=activation if plain else None,
                                   preact=preact))
        self._fusable = (not preact) and isinstance(activation, nn.ReLU)

    def layer_params(self):
        """[(conv, bn-or-None)] per layer, or None if the stack is not conv->bn->relu."""
        if not self._fusable:
            return None
        out = []
        for layer in self.children():
            conv = getattr(layer, "conv", None)
            bn = getattr(layer, "bn", None)
            if conv is None:
                return None
            out.append((conv, bn.bn if bn is not None else None))
        return out


def set_bn_momentum_default(bn_momentum):
    def fn(m):
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)):
            m.momentum = bn_momentum
    return fn


class BNMomentumScheduler(object):
    def __init__(self, model, bn_lambda, last_epoch=-1, setter=set_bn_momentum_default):
        if not isinstance(model, nn.Module):
            raise RuntimeError("Class '{}' is not a PyTorch nn Module".format(type(model).__name__))
        self.model = model
        self.setter = setter
        self.lmbd = bn_lambda
        self.step(last_epoch + 1)
        self.last_epoch = last_epoch

    def step(self, epoch=None):
        if epoch is None:
            epoch = self.last_epoch + 1
        self.last_epoch = epoch
        self.model.apply(self.setter(self.lmbd(epoch)))
