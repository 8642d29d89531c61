"""EdgeConv on the libs2c entry points (s2c_edgeconv_fwd / _bwd) vs a float64 PyTorch evaluation of the reference's
definition (models/graph_module.py:102-109: message = map_edge([x_i, x_j - x_i]); propagate :44-100: add-aggregation
at edge_index[1]), and the fused pointwise heads (voting / proposal) vs float64."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _edgeconv64(x, row, col, mask, W1, b1, W2, b2, relu_mask=None):
    """relu_mask: the kernel's own ReLU decisions (a hidden unit within fp32 rounding of zero may take the other branch;
    one such unit among 2.6 M moves the gradient by ~1e-3 relative, which is a property of ReLU, not of the kernel)."""
    xi, xj = x[col], x[row]
    z = torch.cat([xi, xj - xi], 1)
    pre = F.linear(z, W1, b1)
    h = torch.relu(pre) if relu_mask is None else pre * relu_mask.to(pre.dtype)
    msg = F.linear(h, W2, b2)
    if mask is not None:
        msg = msg * mask.unsqueeze(-1).to(msg.dtype)
    agg = torch.zeros(x.shape[0], W2.shape[0], dtype=x.dtype, device=x.device).index_add(0, col, msg)
    return agg, msg, h


@pytest.mark.parametrize("Nn,E,Cin,Cout,masked,use_agg,use_msg", [
    (2048, 20480, 128, 128, True, True, True),     # the CapNet shape: B=8 x 256 proposals, 10 neighbours each
    (2048, 20480, 128, 128, True, False, True),    # edge_layer: only the messages are used
    (512, 5000, 128, 128, False, True, False),
    (300, 1000, 32, 64, True, True, True),
    (256, 3000, 128, 256, True, True, True),
    (64, 0, 128, 128, False, True, True),          # empty graph
])
def test_edgeconv_matches_float64(Nn, E, Cin, Cout, masked, use_agg, use_msg):
    from scan2cap_b200.models.graph_module import EdgeConv
    torch.manual_seed(Nn + E + Cin)
    conv = EdgeConv(Cin, Cout, "add").to(DEV)
    x = torch.randn(Nn, Cin, device=DEV, requires_grad=True)
    row = torch.randint(0, Nn, (E,), device=DEV)
    col = torch.randint(0, Nn, (E,), device=DEV)
    mask = (torch.rand(E, device=DEV) > 0.3) if masked else None
    out, msg = conv(x, torch.stack([row, col], 0), mask, need_aggregate=use_agg)
    relu_mask = None
    if E > 0:
        saved = msg.grad_fn.saved_tensors          # (row, col, mask8, W1, b1, W2, z, Y1) of _EdgeConvFn
        relu_mask = (saved[7] + saved[4]).detach() > 0      # the decisions the kernels took: relu(Y1 + b1)
    g_out = torch.randn(Nn, Cout, device=DEV)
    g_msg = torch.randn(E, Cout, device=DEV)
    loss = 0
    if use_agg:
        loss = loss + (out * g_out).sum()
    if use_msg:
        loss = loss + (msg * g_msg).sum()
    loss.backward()

    p64 = [p.detach().double().requires_grad_(True) for p in
           (conv.map_edge[0].weight, conv.map_edge[0].bias, conv.map_edge[2].weight, conv.map_edge[2].bias)]
    x64 = x.detach().double().requires_grad_(True)
    agg_r, msg_r, h = _edgeconv64(x64, row, col, mask, *p64, relu_mask=relu_mask)
    if E > 0:
        with torch.no_grad():
            pre = F.linear(torch.cat([x64[col], x64[row] - x64[col]], 1), p64[0], p64[1])
            flipped = (pre > 0) != relu_mask
            assert int(flipped.sum()) <= 50 and float(pre[flipped].abs().max() if flipped.any() else 0.0) < 1e-5
    loss_r = 0
    if use_agg:
        loss_r = loss_r + (agg_r * g_out.double()).sum()
    if use_msg:
        loss_r = loss_r + (msg_r * g_msg.double()).sum()
    loss_r.backward()
    if E == 0:
        assert float(msg.abs().sum()) == 0 and (not use_agg or float(out.abs().max()) == 0)
        assert float(x.grad.abs().max()) == 0 and float(conv.map_edge[0].weight.grad.abs().max()) == 0
        return
    assert _rel(msg, msg_r) < 1e-5
    if use_agg:
        assert _rel(out, agg_r) < 1e-5
    tol = 2e-5
    assert _rel(x.grad, x64.grad) < tol
    for p, q, name in zip((conv.map_edge[0].weight, conv.map_edge[0].bias, conv.map_edge[2].weight, conv.map_edge[2].bias),
                          p64, ("W1", "b1", "W2", "b2")):
        assert _rel(p.grad, q.grad) < tol, name


@pytest.mark.parametrize("aggr", ["mean", "max"])
def test_edgeconv_other_aggregations(aggr):
    """aggr = mean / max (never used by CapNet): the fused call produces the messages, the aggregation is torch."""
    from scan2cap_b200.models.graph_module import EdgeConv
    torch.manual_seed(3)
    Nn, E, C = 200, 1500, 128
    conv = EdgeConv(C, C, aggr).to(DEV)
    x = torch.randn(Nn, C, device=DEV, requires_grad=True)
    row = torch.randint(0, Nn, (E,), device=DEV)
    col = torch.randint(0, Nn, (E,), device=DEV)
    out, msg = conv(x, torch.stack([row, col], 0), None)
    out.sum().backward()
    with torch.no_grad():
        p = [q.double() for q in (conv.map_edge[0].weight, conv.map_edge[0].bias, conv.map_edge[2].weight, conv.map_edge[2].bias)]
        _, msg_r, _ = _edgeconv64(x.detach().double(), row, col, None, *p)
        want = torch.zeros(Nn, C, dtype=torch.float64, device=DEV)
        if aggr == "mean":
            deg = torch.zeros(Nn, dtype=torch.float64, device=DEV).index_add(0, col, torch.ones(E, dtype=torch.float64, device=DEV))
            want = want.index_add(0, col, msg_r) / deg.clamp_min(1).unsqueeze(-1)
        else:
            want = want.fill_(float("-inf")).scatter_reduce(0, col.unsqueeze(-1).expand_as(msg_r), msg_r, "amax")
            want = torch.where(torch.isinf(want), torch.zeros_like(want), want)
    assert _rel(out, want) < 1e-5
    assert x.grad is not None and torch.isfinite(x.grad).all()
