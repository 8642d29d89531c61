"""Fused shared-MLP (+ max-pool) on the tcgen05 kernels of libs2c.so, as one autograd Function.

Arithmetic of the reference's SharedMLP (lib/pointnet2/pytorch_utils.py:11-36, 88-120: [1x1 conv without bias ->
BatchNorm (batch statistics in training) -> ReLU] per layer) followed by the max over nsample
(lib/pointnet2/pointnet2_modules.py:255-257) -- or without pooling (nsample = 1) for PointnetFPModule.

Forward: one tensor-core kernel per layer; only the PRE-BatchNorm output of every layer is written to HBM (it is
also what the backward pass needs), the normalised / rectified activations exist only inside the next kernel's
operand staging; the batch statistics come out of the GEMM epilogue; the last BatchNorm + ReLU is folded into the
pooling kernel.
Backward: BatchNorm's batch-statistics backward is affine per channel in (g, y):
    dY = a*g + b*y + c,  a = gamma*invstd, b = -a*invstd*mean(g*xhat), c = -a*mean(g) - b*mean
so each layer's dY tile is formed on the fly from the stored y and the masked upstream gradient g.
"""
import torch
from torch.autograd import Function

from . import _ext_mlp
from ..linear_simt import gemm

# Test hook (tests/parity_utils.py): when set to a list, every forward call appends the tensors that define its
# discontinuous decisions -- pre-BatchNorm outputs + folded BatchNorm affine of every layer (ReLU masks) and the
# pooling arg-max -- so a parity test can tell "different arithmetic" from "a ReLU / arg-max decision flipped by
# fp32 rounding".  Never set by the product.
CAPTURE = None


def _bn_coefficients(s1, s2, R, bn, training):
    """mean / invstd (float64) of this layer's BatchNorm and the folded fp32 (scale, shift); updates running stats
    (one s2c_bn_finalize launch; semantics of torch's batch_norm: unbiased running variance, momentum blend)."""
    return _ext_mlp.bn_finalize(s1, s2, R, bn, training)


def _fold_conv_bias(bias, bn, training, mean, scale, shift):
    """A conv bias in front of BatchNorm (the Conv1d heads of voting_module.py:27-31 / proposal_module.py:46-50): the
    GEMM runs without it.  Batch statistics: the normalised output does not depend on it, only the running mean does
    (+ momentum * bias); running statistics: it moves the folded shift by scale * bias, and the mean the backward pass
    centres the bias-free output with by -bias."""
    batch = training or not bn.track_running_stats
    if batch:
        if training and bn.track_running_stats:
            with torch.no_grad():
                if bn.momentum is not None:
                    bn.running_mean.add_(bias, alpha=float(bn.momentum))
                else:
                    bn.running_mean.addcdiv_(bias, bn.num_batches_tracked.to(bias.dtype))
        return mean, scale, shift
    return mean - bias.double(), scale, torch.addcmul(shift, scale, bias)


def _first_layer_weight(W, K, lda, xyz_gap):
    """The first layer's weight in the column layout of `rows` -> (W', K').
    xyz_gap: rows are [x, y, z, 0 | K-3 features | zero pad] (lda floats): W' = [W[:, :3], 0, W[:, 3:], 0];
    else rows may be zero-padded to a multiple of 4 columns: W' = [W, 0]."""
    N = W.shape[0]
    if xyz_gap:
        C = K - 3
        parts = [W[:, :3], W.new_zeros((N, 1)), W[:, 3:]]
        if lda > 4 + C:
            parts.append(W.new_zeros((N, lda - 4 - C)))
        return torch.cat(parts, 1), lda
    if K % 4 != 0 and lda >= (K + 3) // 4 * 4:
        k4 = (K + 3) // 4 * 4
        return torch.nn.functional.pad(W, (0, k4 - K)), k4
    return W, K


def _first_layer_width(K, lda, xyz_gap):
    """Column count K' of _first_layer_weight(W, K, lda, xyz_gap) without building it."""
    if xyz_gap:
        return lda
    if K % 4 != 0 and lda >= (K + 3) // 4 * 4:
        return (K + 3) // 4 * 4
    return K


def _first_layer_weight_grad(dWp, K, xyz_gap):
    """inverse column mapping of _first_layer_weight for the gradient."""
    if xyz_gap:
        return torch.cat([dWp[:, :3], dWp[:, 4:4 + K - 3]], 1)
    return dWp[:, :K]


def _input_blocks(K, lda, xyz_gap):
    """Column blocks (col0, width) of the first layer's input whose gradient the tensor-core dgrad kernel can write
    (widths 256/128/64, 16-byte aligned): the feature block of gap-layout rows, or all columns of plain rows."""
    col, left = (4, K - 3) if xyz_gap else (0, K)
    if left <= 0 or left % 64 != 0 or (not xyz_gap and lda != K):
        return None
    blocks = []
    while left > 0:
        w = 256 if left >= 256 else (128 if left >= 128 else 64)
        blocks.append((col, w))
        col += w
        left -= w
    return blocks


def _w0(ctx, W, K, lda, xyz_gap):
    """The first layer's weight in the rows' column layout: the forward pass's copy when it is still there."""
    w0 = getattr(ctx, "w0", None)
    if w0 is not None:
        return w0, w0.shape[1]
    return _first_layer_weight(W, K, lda, xyz_gap)


class _FusedMLPPool(Function):
    @staticmethod
    def forward(ctx, rows, K, G, ns, training, bns, xyz_gap, need_xyz_grad, capture, *params):
        """rows (R, lda) fp32 with K valid columns, R = G*ns; params = (W1, gamma1, beta1, bias1 | None, W2, ...);
        bns = the BatchNorm modules (running statistics / eps / momentum).  Returns pooled (G, C_last)."""
        L = len(bns)
        R = rows.shape[0]
        Ys, coefs = [], []
        A, scale, shift, k = rows, None, None, K
        # ONE zero-fill per call for every accumulator of the forward AND the backward pass: the float64 statistics of
        # all layers (forward: L x (sum, sumsq); backward: (L + 1) x (sum g, sum g*y)) and the weight gradients the
        # tensor-core kernels reduce into with red.global.add
        wshapes = []
        for l in range(L):
            Wl = params[4 * l]
            wshapes.append((Wl.shape[0], _first_layer_width(K, rows.shape[1], xyz_gap) if l == 0 else Wl[0].numel()))
        need_bwd = any(ctx.needs_input_grad)
        ndw = sum(c * p for c, p in wshapes) if need_bwd else 0
        nst = (2 * L + 1 if need_bwd else L) * 512
        zall = torch.zeros(nst + (ndw + 1) // 2, dtype=torch.float64, device=rows.device)
        zstats = zall[:L * 512].view(L, 2, 256)
        ctx.zero_ws = (zall[L * 512:nst].view(L + 1, 2, 256), zall[nst:].view(torch.float32)[:ndw]) if need_bwd else None
        W0 = None
        for l in range(L):
            W = params[4 * l].reshape(params[4 * l].shape[0], -1)
            need_stats = training or not bns[l].track_running_stats
            if l == 0:  # the weight in the column layout of the rows (zero columns where the rows are padding)
                W, k = _first_layer_weight(W, K, A.shape[1], xyz_gap)
                W0 = W.detach()
            res = _ext_mlp.mlp_layer_fwd(A, W, scale, shift, want_stats=need_stats, K=k, stats=zstats[l])
            Y, s1, s2 = res if need_stats else (res, None, None)
            mean, invstd, scale, shift = _bn_coefficients(s1, s2, R, bns[l], training)
            if params[4 * l + 3] is not None:
                mean, scale, shift = _fold_conv_bias(params[4 * l + 3].detach(), bns[l], training, mean, scale, shift)
            Ys.append(Y)
            coefs.append((mean, invstd, scale, shift))
            A, k = Y, W.shape[0]
        pooled, argmax = _ext_mlp.pool_fwd(Ys[-1], G, ns, scale, shift, want_argmax=True)
        if CAPTURE is not None and capture:
            CAPTURE.append(dict(G=G, ns=ns, Ys=list(Ys), affine=[(c[2], c[3]) for c in coefs], argmax=argmax,
                                pooled=pooled))
        ctx.save_for_backward(rows, argmax, *Ys, *[t for c in coefs for t in c], *params)
        ctx.w0 = W0  # the first layer's weight in the rows' column layout (the backward pass multiplies by it again)
        ctx.meta = (K, G, ns, L, bool(training), [bool(training or not b.track_running_stats) for b in bns],
                    bool(xyz_gap), bool(need_xyz_grad))
        return pooled

    @staticmethod
    def backward(ctx, dpool):
        K, G, ns, L, training, batch_stats, xyz_gap, need_xyz_grad = ctx.meta
        saved = ctx.saved_tensors
        rows, argmax = saved[0], saved[1]
        Ys = saved[2:2 + L]
        coefs = [saved[2 + L + 4 * l: 2 + L + 4 * l + 4] for l in range(L)]
        params = saved[2 + 5 * L:]
        R = rows.shape[0]
        grads = [None] * (4 * L)
        dpool = dpool.contiguous()
        grad_rows = None
        wshapes = []
        for l in range(L):
            Wl = params[4 * l].reshape(params[4 * l].shape[0], -1)
            wshapes.append((Wl.shape[0], _first_layer_width(K, rows.shape[1], xyz_gap) if l == 0 else Wl.shape[1]))
        # statistics accumulators and weight-gradient buffers: zero-filled by the forward pass together with its own
        # (a second backward through the same graph allocates fresh ones)
        ws, ctx.zero_ws = getattr(ctx, "zero_ws", None), None
        if ws is not None:
            zstats, zdw = ws
        else:
            zstats = torch.zeros((L + 1, 2, 256), dtype=torch.float64, device=rows.device)
            zdw = torch.zeros(sum(c * p for c, p in wshapes), dtype=torch.float32, device=rows.device)
        woff = [0]
        for c, p in wshapes:
            woff.append(woff[-1] + c * p)
        dw_buf = lambda l: zdw[woff[l]:woff[l + 1]].view(wshapes[l])

        def affine(l, sum_g, sum_gy):
            """BatchNorm backward of layer l as dY = a*g + b*y + c; also its gamma / beta gradients."""
            mean, invstd, _, _ = coefs[l]
            gamma = params[4 * l + 1]
            grads[4 * l + 1], grads[4 * l + 2], a, b, c = _ext_mlp.bn_backward_coeffs(
                sum_g, sum_gy, mean, invstd, gamma, R, batch_stats[l])
            if params[4 * l + 3] is not None:
                # a bias in front of a batch-statistics BatchNorm has an identically zero gradient (sum_r dY = 0);
                # with running statistics dY = a*g and the bias gradient is a * sum_r g
                grads[4 * l + 3] = (torch.zeros_like(params[4 * l + 3]) if batch_stats[l]
                                    else (a.double() * sum_g).to(a.dtype))
            return a, b, c

        # last layer: its masked gradient is the pooled gradient at the arg-max sample -> sums straight from dpool
        l = L - 1
        _, _, sc_l, sh_l = coefs[l]
        sum_g, sum_gy = _ext_mlp.pool_bwd_stats(dpool, argmax, Ys[l], ns, sc_l, sh_l, stats=zstats[L])
        g = None  # dense masked gradient of layer l (None while it is still "pooled")
        while l >= 0:
            W = params[4 * l].reshape(params[4 * l].shape[0], -1)
            a, b, c = affine(l, sum_g, sum_gy)
            Y = Ys[l]
            fused = l > 0 and _ext_mlp.bwd_data_supported(Y.shape[1], Ys[l - 1].shape[1])
            if fused:
                _, _, sc_p, sh_p = coefs[l - 1]
                if g is None:
                    g_prev, dY, sum_g, sum_gy = _ext_mlp.mlp_layer_bwd_data(
                        Y, a, b, c, W, Ys[l - 1], sc_p, sh_p, dpool=dpool, argmax=argmax, ns=ns, last_scale=coefs[l][2],
                        last_shift=coefs[l][3], stats=zstats[l - 1])
                else:
                    g_prev, dY, sum_g, sum_gy = _ext_mlp.mlp_layer_bwd_data(Y, a, b, c, W, Ys[l - 1], sc_p, sh_p, G=g,
                                                                            stats=zstats[l - 1])
                if _ext_mlp.bwd_weight_supported(Y.shape[1], Ys[l - 1].shape[1], dY.stride(0), Ys[l - 1].stride(0)):
                    grads[4 * l] = _ext_mlp.mlp_layer_bwd_weight(dY, Ys[l - 1], Ys[l - 1].shape[1], sc_p, sh_p,
                                                                 out=dw_buf(l)).view_as(params[4 * l])
                elif _ext_mlp.wgrad_blocked_supported(Y.shape[1], Ys[l - 1].shape[1], dY.stride(0), Ys[l - 1].stride(0)):
                    grads[4 * l] = _ext_mlp.mlp_layer_bwd_weight_blocked(dY, Ys[l - 1], Ys[l - 1].shape[1], sc_p,
                                                                        sh_p, out=dw_buf(l)).view_as(params[4 * l])
                else:
                    Xp = torch.relu_(torch.addcmul(sh_p, Ys[l - 1], sc_p))
                    grads[4 * l] = (dY.t() @ Xp).view_as(params[4 * l])
                g = g_prev
            else:
                if g is None:  # materialise the pooled gradient (ReLU mask of the last layer applied)
                    g = torch.zeros((G, ns, Y.shape[1]), dtype=dpool.dtype, device=dpool.device)
                    g.scatter_(1, argmax.long().unsqueeze(1), dpool.unsqueeze(1))
                    g = g.view(R, -1) * (torch.addcmul(coefs[l][3], Y, coefs[l][2]) > 0)
                if l == 0:
                    Wp, Kp = _w0(ctx, W, K, rows.shape[1], xyz_gap)
                    blocks = _input_blocks(K, rows.shape[1], xyz_gap) if ctx.needs_input_grad[0] else None
                    if (not ctx.needs_input_grad[0] and
                            _ext_mlp.bwd_weight_supported(Y.shape[1], Kp, g.stride(0), Y.stride(0), rows.stride(0))):
                        # no gradient w.r.t. the input (SA1): dY is formed inside the weight-gradient kernel
                        dWp = _ext_mlp.mlp_layer_bwd_weight(g, rows, Kp, a=a, b=b, c=c, Y=Y, out=dw_buf(0))
                        grads[0] = _first_layer_weight_grad(dWp, K, xyz_gap).reshape(params[0].shape)
                        break
                    if blocks is not None and _ext_mlp.bwd_data_supported(Y.shape[1], 64):
                        # input gradient on the tensor cores, block by block, straight into the (R, lda) gradient rows
                        grad_rows = torch.empty_like(rows)
                        dY = None
                        for i, (c0, w) in enumerate(blocks):
                            d = _ext_mlp.mlp_layer_bwd_input(g, Y, a, b, c, Wp, c0, w, grad_rows, want_dY=(i == 0))
                            dY = d if i == 0 else dY
                        if xyz_gap:
                            if need_xyz_grad:   # the four coordinate columns: a skinny fp32 GEMM (s2c_gemm)
                                gemm(dY, (Wp.stride(0), Wp.stride(1)), Wp, dY.shape[0], 4, dY.shape[1], out=grad_rows[:, :4])
                            else:
                                grad_rows[:, :4] = 0.0
                            if rows.shape[1] > 4 + K - 3:
                                grad_rows[:, 4 + K - 3:] = 0.0
                        if _ext_mlp.bwd_weight_supported(Y.shape[1], Kp, dY.stride(0), rows.stride(0)):
                            dWp = _ext_mlp.mlp_layer_bwd_weight(dY, rows, Kp, out=dw_buf(0))
                        elif _ext_mlp.wgrad_blocked_supported(Y.shape[1], Kp, dY.stride(0), rows.stride(0)):
                            dWp = _ext_mlp.mlp_layer_bwd_weight_blocked(dY, rows, Kp, out=dw_buf(0))
                        else:
                            dWp = dY.t() @ rows[:, :Kp]
                        grads[0] = _first_layer_weight_grad(dWp, K, xyz_gap).reshape(params[0].shape)
                        break
                dY = torch.addcmul(c, g, a).addcmul_(Y, b)
                if l > 0:
                    _, _, sc_p, sh_p = coefs[l - 1]
                    pre = torch.addcmul(sh_p, Ys[l - 1], sc_p)
                    grads[4 * l] = (dY.t() @ torch.relu(pre)).view_as(params[4 * l])
                    g = (dY @ W) * (pre > 0)
                    sum_g = g.sum(0, dtype=torch.float64)
                    sum_gy = (g * Ys[l - 1]).sum(0, dtype=torch.float64)
                else:
                    Wp, Kp = _w0(ctx, W, K, rows.shape[1], xyz_gap)
                    if _ext_mlp.bwd_weight_supported(Y.shape[1], Kp, dY.stride(0), rows.stride(0)):
                        dWp = _ext_mlp.mlp_layer_bwd_weight(dY, rows, Kp)
                    else:
                        dWp = dY.t() @ rows[:, :Kp]
                    grads[0] = _first_layer_weight_grad(dWp, K, xyz_gap).reshape(params[0].shape)
                    if ctx.needs_input_grad[0]:
                        grad_rows = dY @ Wp
                        if rows.shape[1] != Kp:
                            grad_rows = torch.nn.functional.pad(grad_rows, (0, rows.shape[1] - Kp))
            l -= 1
        return (grad_rows, None, None, None, None, None, None, None, None) + tuple(grads)


def fused_mlp_maxpool(rows, K, G, ns, layers, training, xyz_gap=False, need_xyz_grad=True, capture=True):
    """rows (G*ns, >=K) -> (G, C_last): SharedMLP `layers` = [(conv, bn)] then max over each group's ns rows.
    xyz_gap: the rows are [x, y, z, 0 | K-3 features | zero pad] (the padded layout of the fused query+group kernel);
    need_xyz_grad=False skips the gradient of the three coordinate columns (no grad flows to xyz / new_xyz).
    capture=False keeps the call out of the CAPTURE test hook (the pointwise heads: they are not SA / FP modules)."""
    params, bns = [], []
    for conv, bn in layers:
        assert bn is not None, "fused path: conv followed by BatchNorm"
        params += [conv.weight, bn.weight, bn.bias, conv.bias]
        bns.append(bn)
    return _FusedMLPPool.apply(rows, K, G, ns, training, bns, xyz_gap, need_xyz_grad, capture, *params)


def fusable(layers):
    if layers is None or len(layers) == 0:
        return False
    for conv, bn in layers:
        n = conv.weight.shape[0]
        if bn is None or not bn.affine or n % 16 != 0 or n > 256:
            return False
    return True


# ---------------------------------------------------------------------------------------------------------------------
# Final (bias, no BatchNorm, no ReLU) layer of the pointwise heads -- Conv1d(256, 3+256) of VotingModule
# (models/voting_module.py:31,55), Conv1d(128, 2+3+NH*2+NS*4+NC) of ProposalModule (models/proposal_module.py:52) and
# Linear(128, num_bins+1) of GraphModule.edge_predict (models/graph_module.py:151) -- on the same tensor-core kernels.
# The output width is padded to the kernels' tile widths (blocks of 256 / 128 / 64 columns, zero weight rows).

_CONST = {}


def _const(device, n, value):
    key = (str(device), int(n), float(value))
    t = _CONST.get(key)
    if t is None:
        t = _CONST[key] = torch.full((int(n),), float(value), dtype=torch.float32, device=device)
    return t


def _col_blocks(n):
    """n (multiple of 64) -> [(col0, width)] with widths 256 / 128 / 64."""
    blocks, c = [], 0
    while c < n:
        w = 256 if n - c >= 256 else (128 if n - c >= 128 else 64)
        blocks.append((c, w))
        c += w
    return blocks


class _LinearRows(Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        """x (R, K) fp32 rows (K multiple of 64), weight (N, K), bias (N) or None -> (R, N) = x W^T + b."""
        R, K = x.shape
        N = weight.shape[0]
        Np = (N + 63) // 64 * 64
        Wp = torch.nn.functional.pad(weight, (0, 0, 0, Np - N)) if Np != N else weight.contiguous()
        Y = torch.empty((R, Np), dtype=torch.float32, device=x.device)
        for n0, w in _col_blocks(Np):
            _ext_mlp.mlp_layer_fwd(x, Wp[n0:n0 + w], want_stats=False, out=Y, col0=n0)
        ctx.save_for_backward(x, Wp)
        ctx.N = N
        ctx.has_bias = bias is not None
        out = Y[:, :N]
        return out + bias if bias is not None else (out if Np == N else out.contiguous())

    @staticmethod
    def backward(ctx, dy):
        x, Wp = ctx.saved_tensors
        R, K = x.shape
        N, Np = ctx.N, Wp.shape[0]
        dYp = torch.nn.functional.pad(dy, (0, Np - N)) if Np != N else dy.contiguous()
        db = _ext_mlp.col_sum(dYp)[:N] if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        dx = dW = None
        if ctx.needs_input_grad[0]:
            # dx = dY Wp: the first-layer input-gradient kernel with identity BatchNorm-backward coefficients
            one, zero = _const(x.device, Np, 1.0), _const(x.device, Np, 0.0)
            dx = torch.empty_like(x)
            for c0, w in _col_blocks(K):
                _ext_mlp.mlp_layer_bwd_input(dYp, dYp, one, zero, zero, Wp, c0, w, dx, want_dY=False)
        if ctx.needs_input_grad[1]:
            dWp = torch.zeros((Np, K), dtype=torch.float32, device=x.device)
            for n0, w in _col_blocks(Np):
                step = 256 if w <= 128 else 128
                for c0 in range(0, K, step):
                    _ext_mlp.mlp_layer_bwd_weight(dYp[:, n0:n0 + w], x, min(step, K - c0), out=dWp[n0:n0 + w], col0=c0)
            dW = dWp[:N]
        return dx, dW, db


def linear_rows_supported(x, weight):
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[1] % 64 == 0 and x.stride(1) == 1
            and x.stride(0) % 4 == 0 and weight.dim() == 2 and weight.shape[1] == x.shape[1])


def linear_rows(x, weight, bias=None):
    """x (R, K) @ weight (N, K)^T + bias on the tcgen05 layer kernels (3xTF32), forward and backward."""
    if not linear_rows_supported(x, weight):
        raise RuntimeError("linear_rows: x must be a CUDA fp32 (R, K) matrix with K a multiple of 64 and aligned rows "
                           "(tensor-core kernels of libs2c); got %s / %s" % (tuple(x.shape), tuple(weight.shape)))
    if x.data_ptr() % 16 != 0:
        x = x.contiguous()
    return _LinearRows.apply(x, weight, bias)
