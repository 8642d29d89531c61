import torch, torch.nn as nn, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
dev = "cuda"
rel = lambda u, v: float((u.double() - v.double()).abs().max() / v.double().abs().max())
for (B, M, ns, C, widths) in [(2, 2048, 64, 7, (64, 64, 128)), (2, 1024, 32, 131, (128, 128, 256)), (2, 256, 16, 259, (128, 128, 128))]:
    x = torch.randn(B, C, M, ns, device=dev)
    ws, cin = [], C
    for w in widths:
        ws.append(torch.randn(w, cin, device=dev) * (2.0 / cin) ** 0.5); cin = w
    g = torch.randn(B, widths[-1], M, device=dev)

    def path_conv(dt):
        xx = x.to(dt).requires_grad_(True); W = [w.to(dt).clone().requires_grad_(True) for w in ws]
        h = xx
        for w in W:
            h = F.conv2d(h, w[:, :, None, None]); h = F.batch_norm(h, None, None, torch.ones(w.shape[0], dtype=dt, device=dev), torch.zeros(w.shape[0], dtype=dt, device=dev), True); h = F.relu(h)
        out = F.max_pool2d(h, kernel_size=[1, ns]).squeeze(-1)
        (out * g.to(dt)).sum().backward()
        return [out.detach()] + [w.grad for w in W] + [xx.grad]

    def path_rows(dt):
        xx = x.to(dt).requires_grad_(True); W = [w.to(dt).clone().requires_grad_(True) for w in ws]
        h = xx.permute(0, 2, 3, 1).reshape(-1, C)
        for w in W:
            h = F.linear(h, w); h = F.batch_norm(h, None, None, torch.ones(w.shape[0], dtype=dt, device=dev), torch.zeros(w.shape[0], dtype=dt, device=dev), True); h = F.relu(h)
        out = h.view(B, M, ns, -1).amax(2).transpose(1, 2)
        (out * g.to(dt)).sum().backward()
        return [out.detach()] + [w.grad for w in W] + [xx.grad]
    t = path_conv(torch.float64); t2 = path_rows(torch.float64)
    a = path_conv(torch.float32); b = path_rows(torch.float32)
    print("shape", (B, M, ns, C, widths))
    for i, nm in enumerate(["out", "gW0", "gW1", "gW2", "gx"]):
        print("  %-4s f64 conv-vs-rows %.1e | f32 conv vs truth %.1e | f32 rows vs truth %.1e | f32 conv vs rows %.1e" % (nm, rel(t2[i], t[i]), rel(a[i], t[i]), rel(b[i], t[i]), rel(a[i], b[i])))
# plain GEMM accuracy
A = torch.randn(4096, 256, device=dev); Bm = torch.randn(256, 256, device=dev)
print("linear fp32 vs fp64:", rel(F.linear(A, Bm), F.linear(A.double(), Bm.double())))
print("conv2d fp32 vs fp64:", rel(F.conv2d(A.t().reshape(1, 256, 64, 64), Bm[:, :, None, None]), F.conv2d(A.double().t().reshape(1, 256, 64, 64), Bm.double()[:, :, None, None])))
