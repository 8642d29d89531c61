"""tcgen05 grouped-MLP kernels vs a plain PyTorch float64/float32 reference of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("R,K,N,lda,pro", [
    (1000, 7, 64, 7, False), (4096, 64, 64, 64, True), (128 * 300 + 5, 64, 128, 64, True),
    (5000, 131, 128, 131, False), (3000, 259, 128, 260, True), (20000, 128, 256, 128, True),
    (777, 32, 16, 40, False), (2048 * 64, 8, 64, 8, False),
])
def test_mlp_layer_fwd_matches_float64(R, K, N, lda, pro):
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
    C, s1, s2 = _ext_mlp.mlp_layer_fwd(buf if lda != K else A, W, scale, shift, want_stats=True, K=K)
    torch.cuda.synchronize()
    err = float((C.double() - want).abs().max() / want.abs().max())
    assert err < 2e-6, "3xTF32 GEMM error %g" % err
    # plain fp32 library GEMM for scale: we must be in the same accuracy class
    ref32 = (A64.float() @ W.t())
    err32 = float((ref32.double() - want).abs().max() / want.abs().max())
    assert err < max(4 * err32, 2e-6)
    assert float((s1 - want.sum(0)).abs().max() / want.sum(0).abs().max().clamp_min(1.0)) < 1e-5
    assert float((s2 - (want * want).sum(0)).abs().max() / (want * want).sum(0).abs().max()) < 1e-5
