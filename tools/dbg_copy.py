import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scan2cap_b200.models.backbone_module import padded_point_clouds_like
junk = [torch.full((1 << 24,), float("nan"), device="cuda") for _ in range(8)]
torch.cuda.synchronize(); del junk
for F in (7, 135):
    h = torch.randn(2, 8000, F).pin_memory()
    for nb in (True, False):
        v = padded_point_clouds_like(h.shape, h.dtype, "cuda")
        v.copy_(h, non_blocking=nb)
        torch.cuda.synchronize()
        print("F", F, "non_blocking", nb, "equal", bool(torch.equal(v.cpu(), h)), "nan in view", bool(torch.isnan(v).any()),
              "strides", v.stride(), "ptr%16", v.data_ptr() % 16, "feat ptr%16", v[..., 3:].data_ptr() % 16)
    d = h.cuda()
    v = padded_point_clouds_like(h.shape, h.dtype, "cuda"); v.copy_(d); torch.cuda.synchronize()
    print("F", F, "d2d equal", bool(torch.equal(v, d)))
