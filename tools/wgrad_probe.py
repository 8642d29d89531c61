import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from scan2cap_b200.lib.pointnet2 import _ext_mlp
torch.set_printoptions(precision=4, linewidth=200)
R, C, P = [int(x) for x in sys.argv[1:4]]
mode = sys.argv[4] if len(sys.argv) > 4 else "rand"
torch.manual_seed(0)
if mode == "rand":
    g = torch.randn(R, C, device="cuda"); X = torch.randn(R, P, device="cuda")
else:  # structured: dY[r, m] = 1 if m == m0 ; X[r, n] = n+1 for a single row r0 -> dW[m0, n] = n+1
    g = torch.zeros(R, C, device="cuda"); X = torch.zeros(R, P, device="cuda")
    r0, m0 = 5, 3
    g[r0, m0] = 1.0; X[r0] = torch.arange(1, P + 1, device="cuda").float()
want = g.double().t() @ X.double()
got = _ext_mlp.mlp_layer_bwd_weight(g, X, P)
torch.cuda.synchronize()
err = float((got.double() - want).abs().max() / want.abs().max())
print("R,C,P", R, C, P, "err", err)
if err > 1e-4:
    print("want[:6,:10]\n", want[:6, :10].float())
    print("got[:6,:10]\n", got[:6, :10])
    nz = got.nonzero()
    print("nonzeros of got (first 20):", nz[:20].tolist(), "count", len(nz))
    if mode != "rand":
        print("values at nonzeros:", got[got != 0][:40].tolist())
