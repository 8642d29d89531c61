"""tcgen05 grouped-MLP kernels vs a plain PyTorch float64/float32 reference of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("version", [1, 2])
@pytest.mark.parametrize("R,K,N,lda,pro", [
    (1000, 7, 64, 7, False), (4096, 64, 64, 64, True), (128 * 300 + 5, 64, 128, 64, True),
    (5000, 131, 128, 131, False), (3000, 259, 128, 260, True), (20000, 128, 256, 128, True),
    (777, 32, 16, 40, False), (2048 * 64, 8, 64, 8, False), (5000, 132, 128, 132, True), (3000, 260, 128, 260, True),
    (9000, 512, 256, 512, True), (100, 64, 64, 64, True),
])
def test_mlp_layer_fwd_matches_float64(R, K, N, lda, pro, version):
    from scan2cap_b200.lib.pointnet2 import _ext_mlp
    torch.manual_seed(R + K + N)
    buf = torch.randn(R, lda, device=DEV)
    A = buf[:, :K] if lda != K else buf
    W = torch.randn(N, K, device=DEV) * (2.0 / K) ** 0.5
    scale = shift = None
    A64 = A.double()
    if pro:
        scale = torch.rand(K, device=DEV) + 0.5
        shift = torch.randn(K, device=DEV) * 0.3
        A64 = torch.relu(A64 * scale.double() + shift.double())
    want = A64 @ W.double().t()
    C, s1, s2 = _ext_mlp.mlp_layer_fwd(buf if lda != K else A, W, scale, shift, want_stats=True, K=K, version=version)
    torch.cuda.synchronize()
    err = float((C.double() - want).abs().max() / want.abs().max())
    assert err < 1e-5, "3xTF32 GEMM error %g" % err
    # plain fp32 library GEMM for scale: we must be in the same accuracy class
    ref32 = (A64.float() @ W.t())
    err32 = float((ref32.double() - want).abs().max() / want.abs().max())
    assert err < max(4 * err32, 1e-5)
    assert float((s1 - want.sum(0)).abs().max() / want.sum(0).abs().max().clamp_min(1.0)) < 1e-5
    assert float((s2 - (want * want).sum(0)).abs().max() / (want * want).sum(0).abs().max()) < 1e-5


@pytest.mark.parametrize("B,M,ns,C,widths,training", [
    (2, 256, 16, 7, (64, 64, 128), True), (2, 128, 32, 131, (128, 128, 256), True),
    (1, 64, 16, 259, (128, 128, 128), False), (3, 200, 1, 512, (256, 256), True),
])
def test_fused_mlp_maxpool_matches_torch_path(B, M, ns, C, widths, training):
    """Fused tensor-core stack (forward AND backward, BatchNorm running statistics included) against the
    library-kernel formulation of the same module (F.linear + F.batch_norm + relu + amax)."""
    import copy
    from scan2cap_b200.lib.pointnet2 import pytorch_utils as pt_utils
    from scan2cap_b200.lib.pointnet2.fused_mlp import fused_mlp_maxpool
    from compare_paths import shared_mlp_rows
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(7)
    mlp_a = pt_utils.SharedMLP([C] + list(widths), bn=True).to(DEV)
    for m in mlp_a.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.2)
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
    mlp_b = copy.deepcopy(mlp_a)
    mlp_a.train(training); mlp_b.train(training)
    R = B * M * ns
    x_a = torch.randn(R, C, device=DEV, requires_grad=True)
    x_b = x_a.detach().clone().requires_grad_(True)
    gout = torch.randn(B * M, widths[-1], device=DEV)
    out_a = fused_mlp_maxpool(x_a, C, B * M, ns, mlp_a.layer_params(), training)
    out_b = shared_mlp_rows(x_b, mlp_b.layer_params(), training).view(B * M, ns, -1).amax(1)
    rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12))
    assert rel(out_a, out_b) < 1e-4
    (out_a * gout).sum().backward()
    (out_b * gout).sum().backward()
    # ReLU / max-pool are discontinuous: 1e-6 forward differences flip a few arg-max / sign decisions, each moving
    # single gradient entries by O(1e-3) of the scale -> judge gradients by their relative L2 error
    l2 = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-12))
    # (k flipped arg-max decisions among n pooled entries give a relative L2 difference of about sqrt(k/n);
    #  tools/diag_fused.py shows both paths at 1e-6 of a float64 evaluation when no decision flips)
    assert l2(x_a.grad, x_b.grad) < 5e-3
    for (n, pa), (_, pb) in zip(mlp_a.named_parameters(), mlp_b.named_parameters()):
        assert l2(pa.grad, pb.grad) < 5e-3, n
    for (n, ba), (_, bb) in zip(mlp_a.named_buffers(), mlp_b.named_buffers()):
        assert rel(ba.float(), bb.float()) < 1e-4, n


@pytest.mark.parametrize("R,C,P,ldx,affine,xpro", [
    (4096, 64, 64, 64, False, True), (100000, 128, 64, 64, False, True), (5000, 256, 128, 128, True, True),
    (3333, 128, 259, 260, False, False), (70000, 64, 7, 8, True, False), (2048, 128, 131, 132, True, False),
])
def test_mlp_layer_bwd_weight_matches_float64(R, C, P, ldx, affine, xpro):
    from scan2cap_b200.lib.pointnet2 import _ext_mlp
    torch.manual_seed(R + C + P)
    g = torch.randn(R, C, device=DEV)
    Xbuf = torch.randn(R, ldx, device=DEV)
    X = Xbuf[:, :P]
    d64 = g.double()
    kw = {}
    if affine:
        Y = torch.randn(R, C, device=DEV)
        a, b, c = (torch.randn(C, device=DEV) for _ in range(3))
        d64 = a.double() * g.double() + b.double() * Y.double() + c.double()
        kw.update(a=a, b=b, c=c, Y=Y)
    x64 = X.double()
    if xpro:
        xs, xh = torch.rand(P, device=DEV) + 0.5, torch.randn(P, device=DEV) * 0.3
        x64 = torch.relu(x64 * xs.double() + xh.double())
        kw.update(xs=xs, xh=xh)
    want = d64.t() @ x64
    got = _ext_mlp.mlp_layer_bwd_weight(g, Xbuf, P, **kw)
    torch.cuda.synchronize()
    err = float((got.double() - want).abs().max() / want.abs().max())
    assert err < 2e-5, "wgrad error %g" % err


@pytest.mark.gpu
@pytest.mark.parametrize("training", [True, False])
def test_bn_coefficient_kernels_match_torch_batchnorm(training):
    """s2c_bn_finalize / s2c_bn_backward_coeffs vs nn.BatchNorm1d (forward, running statistics, backward)."""
    from scan2cap_b200.lib.pointnet2 import _ext_mlp
    torch.manual_seed(0)
    R, N = 5000, 96
    y = (torch.randn(R, N, device="cuda") * 3 + 1.5).requires_grad_(True)
    bn_ref = torch.nn.BatchNorm1d(N).cuda()
    bn = torch.nn.BatchNorm1d(N).cuda()
    with torch.no_grad():
        bn_ref.weight.uniform_(0.5, 1.5); bn_ref.bias.uniform_(-1, 1)
        bn_ref.running_mean.uniform_(-1, 1); bn_ref.running_var.uniform_(0.5, 2)
    bn.load_state_dict(bn_ref.state_dict())
    bn_ref.train(training); bn.train(training)
    out_ref = bn_ref(y)
    g = torch.randn_like(out_ref)
    out_ref.backward(g)
    yd = y.detach().double()
    mean, invstd, scale, shift = _ext_mlp.bn_finalize(yd.sum(0), (yd * yd).sum(0), R, bn, training)
    torch.testing.assert_close(torch.addcmul(shift, y.detach(), scale), out_ref.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(bn.running_mean, bn_ref.running_mean, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(bn.running_var, bn_ref.running_var, rtol=1e-6, atol=1e-6)
    assert int(bn.num_batches_tracked) == int(bn_ref.num_batches_tracked)
    gd = g.double()
    gg, gb, a, b, c = _ext_mlp.bn_backward_coeffs(gd.sum(0), (gd * yd).sum(0), mean, invstd, bn.weight, R, training)
    torch.testing.assert_close(gg, bn_ref.weight.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(gb, bn_ref.bias.grad, rtol=1e-4, atol=1e-4)
    dy = a * g + b * y.detach() + c
    torch.testing.assert_close(dy, y.grad, rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("Cf,widths,ns,need_xyz", [(128, (128, 128, 256), 32, True), (256, (128, 128, 128), 16, True),
                                                  (256, (128, 128, 256), 16, False), (4, (64, 64, 128), 64, True)])
def test_fused_mlp_gap_layout_matches_torch_path(Cf, widths, ns, need_xyz):
    """Rows in the padded layout of the fused query+group kernel, [x, y, z, 0 | features | pad]: forward, and the
    backward with the first layer's input gradient on the tensor-core kernel (s2c_mlp_layer_bwd_input)."""
    import copy
    from scan2cap_b200.lib.pointnet2 import pytorch_utils as pt_utils
    from scan2cap_b200.lib.pointnet2.fused_mlp import fused_mlp_maxpool
    from compare_paths import shared_mlp_rows
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(11)
    G = 96
    R = G * ns
    mlp_a = pt_utils.SharedMLP([3 + Cf] + list(widths), bn=True).to(DEV)
    mlp_b = copy.deepcopy(mlp_a)
    x_a = torch.randn(R, 3 + Cf, device=DEV, requires_grad=True)
    x_b = x_a.detach().clone().requires_grad_(True)
    Cp = 4 + (Cf + 3) // 4 * 4
    rows = torch.cat([x_a[:, :3], x_a.new_zeros(R, 1), x_a[:, 3:], x_a.new_zeros(R, Cp - 4 - Cf)], 1)
    gout = torch.randn(G, widths[-1], device=DEV)
    out_a = fused_mlp_maxpool(rows, 3 + Cf, G, ns, mlp_a.layer_params(), True, xyz_gap=True, need_xyz_grad=need_xyz)
    out_b = shared_mlp_rows(x_b, mlp_b.layer_params(), True).view(G, ns, -1).amax(1)
    rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12))
    assert rel(out_a, out_b) < 1e-4
    (out_a * gout).sum().backward()
    (out_b * gout).sum().backward()
    l2 = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-12))
    assert l2(x_a.grad[:, 3:], x_b.grad[:, 3:]) < 5e-3
    if need_xyz:
        assert l2(x_a.grad[:, :3], x_b.grad[:, :3]) < 5e-3
    else:
        assert float(x_a.grad[:, :3].abs().max()) == 0.0
    for (n, pa), (_, pb) in zip(mlp_a.named_parameters(), mlp_b.named_parameters()):
        assert l2(pa.grad, pb.grad) < 5e-3, n


@pytest.mark.gpu
@pytest.mark.parametrize("R,IN,OUT,bias", [(8192, 256, 259, True), (2048, 128, 97, True), (20480, 128, 7, True),
                                           (4096, 512, 256, False), (3000, 64, 8, True), (100, 128, 128, True)])
def test_linear_rows_matches_float64(R, IN, OUT, bias):
    """fused_mlp.linear_rows (final layers of the vote / proposal heads, edge_predict): forward, input gradient, weight
    gradient on the tensor-core kernels, column-sum bias gradient; output widths padded to the kernels' tile widths."""
    from scan2cap_b200.lib.pointnet2 import fused_mlp
    torch.manual_seed(R + IN)
    x = torch.randn(R, IN, device=DEV, requires_grad=True)
    lin = torch.nn.Linear(IN, OUT, bias=bias).to(DEV)
    g = torch.randn(R, OUT, device=DEV)
    y = fused_mlp.linear_rows(x, lin.weight, lin.bias)
    assert y.shape == (R, OUT)
    y.backward(g)
    xd = x.detach().double().requires_grad_(True)
    lind = torch.nn.Linear(IN, OUT, bias=bias).to(DEV).double()
    lind.load_state_dict({k: v.double() for k, v in lin.state_dict().items()})
    lind(xd).backward(g.double())
    rel = lambda a, b: float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-12))
    assert rel(y.detach(), lind(xd).detach()) < 1e-5
    assert rel(x.grad, xd.grad) < 1e-5
    assert rel(lin.weight.grad, lind.weight.grad) < 2e-5
    if bias:
        assert rel(lin.bias.grad, lind.bias.grad) < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("lead,IN,OUT,bias,sliced", [((8, 27), 300, 300, False, True), ((8,), 128, 300, True, True),
                                                     ((8, 256), 128, 512, False, False), ((8, 27), 512, 3500, True, False),
                                                     ((3, 5), 7, 9, True, False), ((1,), 64, 64, True, False)])
def test_linear_simt_matches_float64(lead, IN, OUT, bias, sliced):
    """lib/linear_simt.linear (the caption module's Linear layers on s2c_gemm / s2c_gemm_tn): forward, input gradient,
    weight and bias gradients vs float64, with a strided input view and -- as for the map_topdown terms -- a weight
    that is a column slice of a wider matrix."""
    from scan2cap_b200.lib.linear_simt import linear
    torch.manual_seed(IN + OUT)
    full = torch.randn(*lead[:-1], lead[-1] + 3, IN, device=DEV) if len(lead) > 1 else torch.randn(lead[0] + 3, IN, device=DEV)
    x = (full[..., :lead[-1], :] if len(lead) > 1 else full[:lead[0]]).detach().requires_grad_(True)
    wide = torch.randn(OUT, IN + (40 if sliced else 0), device=DEV, requires_grad=True)
    b = torch.randn(OUT, device=DEV, requires_grad=True) if bias else None
    w = wide[:, 17:17 + IN] if sliced else wide
    y = linear(x, w, b)
    assert y.shape == tuple(lead) + (OUT,)
    g = torch.randn_like(y)
    y.backward(g)
    xd = x.detach().double().requires_grad_(True)
    wd = wide.detach().double().requires_grad_(True)
    bd = b.detach().double().requires_grad_(True) if bias else None
    yd = torch.nn.functional.linear(xd, wd[:, 17:17 + IN] if sliced else wd, bd)
    yd.backward(g.double())
    rel = lambda a, c: float((a.double() - c).abs().max() / c.abs().max().clamp_min(1e-12))
    assert rel(y.detach(), yd.detach()) < 1e-5
    assert rel(x.grad, xd.grad) < 1e-5
    assert rel(wide.grad, wd.grad) < 1e-5
    if bias:
        assert rel(b.grad, bd.grad) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("R,C,NOUT,conv_bias", [(8192, 256, 259, True), (2048, 128, 97, False)])
def test_pointwise_head_matches_float64(R, C, NOUT, conv_bias, training):
    """The Conv1d heads (voting_module.py:27-60 with conv biases in front of BatchNorm; proposal_module.py:46-78 without)
    on the fused kernels vs a float64 evaluation with nn.Conv1d / nn.BatchNorm1d: outputs, running statistics, all
    parameter gradients (the bias in front of a batch-statistics BatchNorm has an exactly zero gradient)."""
    from scan2cap_b200.lib.pointnet2 import fused_mlp
    torch.manual_seed(R + NOUT)
    mk = lambda: (torch.nn.Conv1d(C, C, 1, bias=conv_bias), torch.nn.BatchNorm1d(C), torch.nn.Conv1d(C, C, 1, bias=conv_bias),
                  torch.nn.BatchNorm1d(C), torch.nn.Conv1d(C, NOUT, 1))
    ours = torch.nn.ModuleList(mk()).to(DEV)
    for m in ours:
        if isinstance(m, torch.nn.BatchNorm1d):
            m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.2)
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
    ref = torch.nn.ModuleList(mk()).to(DEV).double()
    ref.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in ours.state_dict().items()})
    ours.train(training); ref.train(training)
    x = torch.randn(R, C, device=DEV, requires_grad=True)
    g = torch.randn(R, NOUT, device=DEV)
    fused_mlp.CAPTURE = []
    try:
        h = fused_mlp.fused_mlp_maxpool(x, C, R, 1, [(ours[0], ours[1]), (ours[2], ours[3])], training)
        cap = fused_mlp.CAPTURE[0]
    finally:
        fused_mlp.CAPTURE = None
    y = fused_mlp.linear_rows(h, ours[4].weight.view(NOUT, C), ours[4].bias)
    y.backward(g)
    # the kernels' own ReLU decisions (an activation within fp32 rounding of zero may take the other branch; one such
    # unit among 2 M moves the gradients by ~1e-3 relative -- a property of ReLU, not of the kernels)
    masks = [(Y * sc + sh) > 0 for Y, (sc, sh) in zip(cap["Ys"], cap["affine"])]
    xd = x.detach().double().requires_grad_(True)
    t = xd.t().unsqueeze(0)                                   # (1, C, R)
    pre1 = ref[1](ref[0](t))
    t = pre1 * masks[0].t().unsqueeze(0).double()
    pre2 = ref[3](ref[2](t))
    t = pre2 * masks[1].t().unsqueeze(0).double()
    for pre, m in ((pre1, masks[0]), (pre2, masks[1])):
        flipped = (pre.detach().squeeze(0).t() > 0) != m
        assert int(flipped.sum()) <= 50 and float(pre.detach().squeeze(0).t()[flipped].abs().max() if flipped.any() else 0.0) < 1e-5
    yd = ref[4](t).squeeze(0).t()
    yd.backward(g.double())
    rel = lambda a, b: float((a.double() - b).norm() / b.norm().clamp_min(1e-30))
    assert rel(y.detach(), yd.detach()) < 1e-5
    assert rel(x.grad, xd.grad) < 1e-4
    for (n, po), (_, pr) in zip(ours.named_parameters(), ref.named_parameters()):
        if conv_bias and training and n in ("0.bias", "2.bias"):
            # exactly zero in exact arithmetic; the float64 evaluation leaves rounding noise
            scale = float(ref[4].bias.grad.norm())
            assert float(po.grad.abs().max()) == 0.0 and float(pr.grad.norm()) < 1e-9 * scale, n
            continue
        assert rel(po.grad, pr.grad) < 1e-4, n
    for (n, bo), (_, br) in zip(ours.named_buffers(), ref.named_buffers()):
        if bo.is_floating_point():
            assert rel(bo, br) < 1e-5, n
        else:
            assert torch.equal(bo, br), n
