"""Stage timing inside the caption cluster kernel (forward): %globaltimer stamps of CTA 0 per word.
slots: 0 step start | 1 S1 done (before barrier) | 2 after barrier | 3 S2 (GRU1) done | 4 after S3+barrier |
5 attention done | 6 after S4 barrier (start GRU2); next word's slot 0 = after S5 + barrier."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn as nn
from scan2cap_b200.lib import caption_decoder as cd
torch.manual_seed(0)
dev = "cuda:0"
B, T, K, E, H, F = 8, 26, 256, 300, 512, 128
mk = lambda *s: torch.randn(*s, device=dev) * 0.3
pre_word, pre_tgt, mapped, obj = mk(B, T, E), mk(B, E), mk(B, K, H), mk(B, K, F)
valid = torch.zeros(B, K, device=dev)
for b in range(B):
    valid[b, torch.randperm(K, device=dev)[:11]] = 1
w_td = mk(E, E + H + F)
c1, c2 = nn.GRUCell(E, H).to(dev), nn.GRUCell(E, H).to(dev)
mh, at, ml = nn.Linear(H, H, bias=False).to(dev), nn.Linear(H, 1, bias=False).to(dev), nn.Linear(F + H, E).to(dev)
run = lambda: cd.topdown_decode(pre_word, pre_tgt, mapped, obj, valid, w_td[:, E:E + H], c1, mh, at, ml, c2)
for _ in range(3):
    run()
torch.cuda.synchronize()
cd.DEBUG_TS = []
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); run(); b.record(); torch.cuda.synchronize()
ts = cd.DEBUG_TS[0].cpu().double()
print("kernel+launch: %.1f us for T=%d" % (a.elapsed_time(b) * 1e3, T))
d = torch.zeros(T - 1, 8)
for s in range(7):
    nxt = ts[:-1, s + 1] if s < 6 else ts[1:, 0]
    d[:, s] = (nxt - ts[:-1, s]) / 1e3
names = ["S1 load+gemv", "barrier1", "S2 GRU1", "barrier2+S3+barrier3", "attention", "lang gemv+barrier4", "S5 GRU2+barrier5"]
for s in range(7):
    print("%-24s median %.2f us" % (names[s], float(d[:, s].median())))
print("per word: %.2f us" % float(((ts[1:, 0] - ts[:-1, 0]) / 1e3).median()))
