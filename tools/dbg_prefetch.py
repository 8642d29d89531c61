import copy, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_engine_gpu as T
from scan2cap_b200.engine import TrainStep
model, DC, batches = T._setup()
if os.environ.get("DBG_POISON", "1") == "1":   # expose reads of uninitialised memory: recycled blocks hold NaN
    junk = [torch.full((1 << 26,), float("nan"), device="cuda") for _ in range(8)]
    junk += [torch.full((n,), float("nan"), device="cuda") for n in (1 << 20, 1 << 16, 1 << 12, 1 << 18, 1 << 22)] * 8
    torch.cuda.synchronize()
    del junk
import scan2cap_b200.engine as E
if os.environ.get("DBG_NOPAD") == "1":
    E.padded_point_clouds_like = lambda shape, dtype, dev: torch.empty(shape, dtype=dtype, device=dev)
eng = TrainStep(model, DC, use_cuda_graph=True, prefetch_indices=os.environ.get("DBG_NOIDX") != "1",
                word_bucket=int(os.environ.get("DBG_BUCKET", "4")), **T.FLAGS)
if os.environ.get("DBG_NOPREFETCH") == "1":
    eng.prefetch = lambda d: None
src = {k: v.to("cuda") for k, v in batches[0].items()} if os.environ.get("DBG_DEV", "1") == "1" else batches[0]
nxt = dict(src)
eng.prefetch(nxt)
for step in range(4):
    cur = nxt
    loss = eng.run(cur)
    torch.cuda.synchronize()
    sig = eng._signature(cur if "num_words" in cur else dict(cur, num_words=eng._words(cur)))
    static = eng._graphs[list(eng._graphs)[0]][0]
    ok, rng = [], []
    if "fps_precomputed" in static:
        fresh = model.backbone_net.sample_indices(static["point_clouds"][..., :3])
        ok = [bool(torch.equal(a[0], b[0])) for a, b in zip(static["fps_precomputed"], fresh)]
        rng = [(int(a[0].min()), int(a[0].max())) for a in static["fps_precomputed"]]
    print("step", step, "loss", float(loss), "fps equal", ok, "ranges", rng,
          "pc nan", bool(torch.isnan(static["point_clouds"]).any()), flush=True)
    nxt = dict(src)
    eng.prefetch(nxt)
