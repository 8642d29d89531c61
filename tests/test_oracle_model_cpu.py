"""CPU test: the oracle restatement of the reference's Python layers (oracle/ref_model.py, ref_loss.py)
reproduces the golden vectors that the reference's OWN modules produced (tests/golden/ref_python_capnet.npz,
written by oracle/validate_vs_reference.py in the build container, where /root/reference is importable)."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_loss as RL
from oracle import ref_model as R
from oracle.validate_vs_reference import CFG, KEYS_F, KEYS_I
from scan2cap_b200 import synthetic
from scan2cap_b200.data.scannet.model_util_scannet import ScannetDatasetConfig

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_python_capnet.npz"))
CASES = {"center": (2, 4000, "center", 42), "corner": (1, 3000, "corner", 7)}


def _sample(v):
    v = v.detach().numpy()
    return v.reshape(-1)[::max(1, v.size // 2000)] if v.size > 4000 else v


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_model_matches_reference_python(case):
    B, N, qm, seed = CASES[case]
    DC = ScannetDatasetConfig()
    V = 120
    vocab, emb, _ = synthetic.make_vocabulary(V)
    R.set_backend(None)
    old = R.NORMALIZE_BY_RECIPROCAL
    R.NORMALIZE_BY_RECIPROCAL = False  # the golden run was the reference on CPU (torch divides there)
    try:
        torch.manual_seed(seed)
        # same construction order as the reference CapNet => same random initial weights
        model = R.CapNet(DC.num_class, vocab, emb, DC.num_heading_bin, DC.num_size_cluster, DC.mean_size_arr,
                         **dict(CFG, query_mode=qm))
        state = {k: v.clone() for k, v in model.state_dict().items()}
        data = synthetic.make_data_dict(B, N, use_normal=True, num_vocabs=V, seed=seed)
        with torch.no_grad():
            probe = model({k: torch.from_numpy(v.copy()) for k, v in data.items()})
        data["ref_box_corner_label"] = probe["bbox_corner"][:, 5].numpy().copy()
        data["ref_box_corner_label"][-1] += 50.0
        model.load_state_dict(state)
        model.train()
        out = model({k: torch.from_numpy(v.copy()) for k, v in data.items()})
        out = RL.get_scene_cap_loss(out, "cpu", DC, None, True, True, True, True)
        out["loss"].backward()
    finally:
        R.NORMALIZE_BY_RECIPROCAL = old
    for k in KEYS_I:
        np.testing.assert_array_equal(_sample(out[k].long()), GOLD["%s/%s" % (case, k)].astype(np.int64), err_msg=k)
    for k in KEYS_F:
        want = GOLD["%s/%s" % (case, k)]
        got = _sample(out[k])
        scale = np.abs(want).max() + 1e-12
        assert np.abs(got.astype(np.float64) - want).max() / scale < 1e-5, k
    names = sorted(n for n, p in model.named_parameters() if p.grad is not None)
    norms = np.array([float(dict(model.named_parameters())[n].grad.double().norm()) for n in names])
    want = GOLD["%s/grad_norms" % case]
    assert len(norms) == len(want)
    assert np.abs(norms - want).max() / want.max() < 1e-4
