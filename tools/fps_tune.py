import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, %r)
from scan2cap_b200.lib.pointnet2 import _ext
from scan2cap_b200 import synthetic
pc,_ = synthetic.make_point_clouds(8, 40000, use_height=False, seed=42)
x = torch.from_numpy(pc[..., :3].copy()).cuda()
for _ in range(3): _ext.furthest_point_sampling(x, 2048)
torch.cuda.synchronize()
ts=[]
for _ in range(10):
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); _ext.furthest_point_sampling(x, 2048); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
print(json.dumps(dict(ms=float(np.median(ts)))))
''' % ROOT
for cl in (2, 4, 8, 16):
    for th in (0, 512):
        env = dict(os.environ, S2C_FPS_CLUSTER=str(cl), S2C_FPS_THREADS=str(th))
        try:
            out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
            print("cluster", cl, "threads", th or 1024, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)
        except Exception as e:
            print("cluster", cl, "threads", th, "failed", e, flush=True)
