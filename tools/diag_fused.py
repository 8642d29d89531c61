import os, sys, copy
os.environ["S2C_FUSED_MLP"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from scan2cap_b200.lib.pointnet2 import pytorch_utils as pt_utils
from scan2cap_b200.lib.pointnet2.fused_mlp import fused_mlp_maxpool
from compare_paths import shared_mlp_rows
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"
l2 = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
for (B, M, ns, C, widths) in [(2, 256, 16, 7, (64, 64, 128)), (2, 128, 32, 131, (128, 128, 256)), (2, 128, 32, 132, (128, 128, 256)), (8, 2048, 64, 8, (64, 64, 128))]:
    torch.manual_seed(7)
    mlp_a = pt_utils.SharedMLP([C] + list(widths), bn=True).to(DEV)
    for m in mlp_a.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.2)
    mlp_b = copy.deepcopy(mlp_a); mlp_c = copy.deepcopy(mlp_a).double()
    R = B * M * ns
    x = torch.randn(R, C, device=DEV)
    gout = torch.randn(B * M, widths[-1], device=DEV)
    xa = x.clone().requires_grad_(True); xb = x.clone().requires_grad_(True); xc = x.double().requires_grad_(True)
    oa = fused_mlp_maxpool(xa, C, B * M, ns, mlp_a.layer_params(), True)
    ob = shared_mlp_rows(xb, mlp_b.layer_params(), True).view(B * M, ns, -1).amax(1)
    oc = shared_mlp_rows(xc, mlp_c.layer_params(), True).view(B * M, ns, -1).amax(1)
    (oa * gout).sum().backward(); (ob * gout).sum().backward(); (oc * gout.double()).sum().backward()
    print("shape", (B, M, ns, C, widths))
    print("  out   fused-vs-f64 %.2e   torch32-vs-f64 %.2e" % (l2(oa, oc), l2(ob, oc)))
    print("  dx    fused-vs-f64 %.2e   torch32-vs-f64 %.2e   fused-vs-torch32 %.2e" % (l2(xa.grad, xc.grad), l2(xb.grad, xc.grad), l2(xa.grad, xb.grad)))
    for (n, pa), (_, pb), (_, pc) in zip(mlp_a.named_parameters(), mlp_b.named_parameters(), mlp_c.named_parameters()):
        print("  %-22s fused-vs-f64 %.2e   torch32-vs-f64 %.2e" % (n, l2(pa.grad, pc.grad), l2(pb.grad, pc.grad)))
