import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from scan2cap_b200.lib.pointnet2 import _ext_mlp
R, K, N = [int(x) for x in sys.argv[1:4]]
torch.manual_seed(0)
A = torch.randn(R, K, device="cuda"); W = torch.randn(N, K, device="cuda") * (2.0 / K) ** 0.5
sc = torch.rand(K, device="cuda") + 0.5; sh = torch.randn(K, device="cuda") * 0.3
C, s1, s2 = _ext_mlp.mlp_layer_fwd(A, W, sc, sh, want_stats=True, version=2)
torch.cuda.synchronize()
want = torch.relu(A.double() * sc.double() + sh.double()) @ W.double().t()
print("R,K,N", R, K, N, "RS", os.environ.get("S2C_MLP_RS"), "err", float((C.double() - want).abs().max() / want.abs().max()), flush=True)
