"""CPU checks against the reference's OWN Python modules (the verbatim copy under baseline/_ref, installed by
baseline/install_ref.py in the build container; skipped where it is absent): run in a subprocess because the CPU
shims neutralise Tensor.cuda process-wide."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "models", "capnet.py"))

SCRIPT = r'''
import json, sys
import numpy as np, torch
sys.path.insert(0, %(root)r)
from baseline import shims, reference_arm as RA
shims.install(ext=None, cpu=True)
from lib.loss_helper_pretrained import get_loss as ref_loss
from models.capnet_pretrained import CapNet as RefCapNet
syn = RA.load_synthetic()
V = 80
vocab, emb, _ = syn.make_vocabulary(V)
torch.manual_seed(0)
ref = RefCapNet("votenet", vocab, emb, use_topdown=True, num_locals=10, query_mode="center", graph_mode="edge_conv",
                num_graph_steps=2, use_relation=True, use_orientation=True)
d = syn.make_pretrained_data_dict(2, num_proposals=256, num_valid=64, num_vocabs=V, seed=3, lang_len=12)
d["ref_box_corner_label"][1] += 50.0            # scene 1: no good box -> masked out of the caption loss
data = {k: torch.from_numpy(v) for k, v in d.items()}
out = ref_loss(ref({k: v.clone() for k, v in data.items()}), mode="votenet", orientation=True)
# the product's loss mirror (pure PyTorch, runs on CPU) on the SAME model outputs
from scan2cap_b200.lib.loss_helper_pretrained import get_loss as our_loss
mine = our_loss({k: (v.detach().clone() if isinstance(v, torch.Tensor) else v) for k, v in out.items()},
                mode="votenet", orientation=True)
res = {k: [float(out[k]), float(mine[k])] for k in ("loss", "cap_loss", "ori_loss", "cap_acc", "ori_acc")}
res["edges"] = [int(x) for x in out["num_edge_source"]]
print("RESULT " + json.dumps(res))
'''


@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref not installed (build container only)")
def test_pretrained_loss_mirror_matches_reference_loss():
    """scan2cap_b200/lib/loss_helper_pretrained.py == the reference's lib/loss_helper_pretrained.py:167-214 on the
    outputs of the reference's own capnet_pretrained model (BASELINE configs[0]) -- every loss / accuracy key."""
    p = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
    assert line, p.stderr[-2000:]
    res = json.loads(line[-1][7:])
    assert res["edges"] == [64, 64]
    for k in ("loss", "cap_loss", "ori_loss", "cap_acc", "ori_acc"):
        a, b = res[k]
        assert abs(a - b) <= 1e-6 * max(1.0, abs(a)), (k, a, b)


@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref not installed (build container only)")
def test_reference_arm_process_is_clean():
    """bench.py --impl reference must map neither libs2c.so nor the oracle restatements (caveat 1)."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ref-device", "cpu",
                        "--no-cpu-baseline"], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference"
