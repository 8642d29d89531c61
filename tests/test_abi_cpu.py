"""No-GPU checks of the drop-in boundary: libs2c.so loads, exports every symbol include/s2c.h declares (and
nothing is declared twice), argument validation returns error codes + messages without touching a device, the
drop-in aliasing works, and the product package never imports the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "s2c.h")).read()
    return re.findall(r"S2C_API\s+(?:const\s+char\s*\*|int|long\s+long)\s*(s2c_\w+)\s*\(", txt)


def test_library_exports_every_declared_symbol():
    from scan2cap_b200 import _lib
    names = _declared()
    assert len(names) == len(set(names)) and len(names) >= 15
    for n in names:
        assert hasattr(_lib.LIB, n), n
    # and every symbol the Python side binds is declared in the header
    for n in _lib.SIGNATURES:
        assert n in names, "%s bound in _lib.py but not declared in include/s2c.h" % n
    assert _lib.LIB.s2c_version() >= 100


def test_argument_validation_without_a_device():
    from scan2cap_b200 import _lib
    L = _lib.LIB
    # nsample out of range -> S2C_ERR_INVALID_ARGUMENT before any CUDA call
    rc = L.s2c_ball_query(None, None, 1, 10, 4, ctypes.c_float(0.1), 0, None, None, None)
    assert rc == 1 and b"nsample" in L.s2c_last_error()
    rc = L.s2c_furthest_point_sampling(None, 1, 0, 4, None, None, None)
    assert rc == 1
    rc = L.s2c_mlp_layer_fwd(None, 8, 128, 8, None, None, None, 24, None, 24, None, None, None)
    assert rc == 1 and b"multiple of 16" in L.s2c_last_error()
    # empty problems are a no-op, not an error
    assert L.s2c_ball_query(None, None, 0, 0, 0, ctypes.c_float(0.1), 4, None, None, None) == 0
    with pytest.raises(_lib.S2CError):
        _lib.call("s2c_knn_adjacency", None, None, None, 1, 2000, 1, 10, 1, 0, 0.5, None, None, None)


def test_argument_validation_of_the_fused_entry_points():
    """The entry points added for the fused training path reject bad shapes with a message, before any CUDA call."""
    from scan2cap_b200 import _lib
    L = _lib.LIB
    P = _lib.CaptionParams()
    P.B, P.T, P.K, P.E, P.H, P.F = 8, 5, 256, 300, 500, 128      # H not a multiple of 64
    P.ld_tdh = 940
    assert L.s2c_caption_decode_fwd(ctypes.byref(P), None) == 1 and b"multiple of 64" in L.s2c_last_error()
    P.H = 512
    assert L.s2c_caption_decode_bwd(ctypes.byref(P), None) == 1 and b"null" in L.s2c_last_error()
    assert L.s2c_bn_finalize(None, None, 0, 64, None, None, 1e-5, 0.1, 1, 1, None, None, None, None, None, None, None, None) == 1
    assert L.s2c_bn_backward_coeffs(None, None, None, None, None, 10, 0, 1, None, None, None, None, None, None) == 0   # N = 0: no-op
    assert L.s2c_group_rows_grad(None, 4, 3, 8, None, 1, 10, 5, ctypes.c_float(1.0), None, None) == 1                    # ld < c0 + C
    assert L.s2c_mlp_layer_bwd_input(None, 64, None, 64, 100, 64, None, None, None, None, 132, 96, None, 132, None, None, None) == 1
    assert b"64, 128 or 256" in L.s2c_last_error()
    assert L.s2c_gemm_tn(None, 8, None, 8, 4, 0, 8, None, 8, None, None) == 0                                           # M = 0: no-op
    assert L.s2c_col_sum(None, 4, 10, 8, None, None) == 1                                                                # lda < M
    assert L.s2c_ball_query_grid_workspace_bytes(8, 40000) > 8 * 40000 * 16
    # round-2 entry points: EdgeConv, split grid build / prebuilt query, flat Adam
    assert L.s2c_edgeconv_workspace_bytes(20480, 128, 128, 1) > 20480 * (128 + 128 + 256) * 4
    assert L.s2c_edgeconv_workspace_bytes(20480, 128, 128, 0) < 4 * 1024 * 1024
    rc = L.s2c_edgeconv_fwd(None, 10, 128, None, None, None, 5, None, None, None, None, 100, None, None, None, None, None, None)
    assert rc == 1 and b"64, 128 or 256" in L.s2c_last_error()
    rc = L.s2c_edgeconv_fwd(None, 10, 100, None, None, None, 5, None, None, None, None, 128, None, None, None, None, None, None)
    assert rc == 1 and b"multiple of 64" in L.s2c_last_error()
    rc = L.s2c_edgeconv_bwd(None, None, 10, 128, None, None, None, 5, None, None, None, 128, None, None, None, None, None, None,
                            None, None, None)
    assert rc == 1 and b"null" in L.s2c_last_error()
    assert L.s2c_ball_query_grid_build(None, 2, 5000, ctypes.c_float(-1.0), None, 0, None) == 1 and b"radius" in L.s2c_last_error()
    assert L.s2c_ball_query_grid_build(None, 0, 5000, ctypes.c_float(0.2), None, 0, None) == 0                           # B = 0: no-op
    rc = L.s2c_query_and_group_grid_prebuilt(None, None, None, 1, 5000, 16, 0, 0, 0, ctypes.c_float(0.2), 0, 0, 0, None, None,
                                             None, 0, None)
    assert rc == 1 and b"nsample" in L.s2c_last_error()
    assert L.s2c_adam_step(None, None, None, None, 10, None, None, 1, None, ctypes.c_float(1.0), None) == 1              # n % 4 != 0
    assert b"multiple of 4" in L.s2c_last_error()
    assert L.s2c_adam_step(None, None, None, None, 0, None, None, 1, None, ctypes.c_float(1.0), None) == 0               # n = 0: no-op
    # fp32 GEMM of the caption Linear layers, pipeline probe of the layer kernels
    assert L.s2c_gemm(None, 8, 1, None, 1, 8, None, 0, 4, 16, 8, None, 8, None) == 1 and b"bad sizes" in L.s2c_last_error()   # ldc < N
    assert L.s2c_gemm(None, 8, 1, None, 1, 8, None, 0, 0, 16, 8, None, 16, None) == 0                                    # M = 0: no-op
    assert L.s2c_gemm(None, 8, 1, None, 1, 8, None, 0, 4, 16, 8, None, 16, None) == 1 and b"null" in L.s2c_last_error()
    assert L.s2c_mlp_probe(None, 0) == 0                                                                                # switches it off


def test_first_layer_weight_layouts_round_trip():
    """Host logic of the fused MLP: the first layer's weight in the column layout of the grouped rows
    ([x,y,z,0 | features | pad], or zero-padded plain rows) and the inverse mapping of its gradient."""
    import torch
    from scan2cap_b200.lib.pointnet2 import fused_mlp as fm
    W = torch.arange(6 * 131, dtype=torch.float32).view(6, 131)          # SA2: 3 + 128 input channels
    Wp, Kp = fm._first_layer_weight(W, 131, 132, True)
    assert Kp == 132 and Wp.shape == (6, 132)
    assert torch.equal(Wp[:, :3], W[:, :3]) and float(Wp[:, 3].abs().sum()) == 0 and torch.equal(Wp[:, 4:], W[:, 3:])
    assert torch.equal(fm._first_layer_weight_grad(Wp, 131, True), W)
    W7 = torch.randn(4, 7)                                               # SA1 with C=4: rows of 8 floats
    Wp, Kp = fm._first_layer_weight(W7, 7, 8, True)
    assert Kp == 8 and torch.equal(fm._first_layer_weight_grad(Wp, 7, True), W7) and float(Wp[:, 3].abs().sum()) == 0
    Wq, Kq = fm._first_layer_weight(W7, 7, 8, False)                     # plain rows padded to a multiple of 4
    assert Kq == 8 and torch.equal(Wq[:, :7], W7) and torch.equal(fm._first_layer_weight_grad(Wq, 7, False), W7)
    assert fm._input_blocks(131, 132, True) == [(4, 128)]
    assert fm._input_blocks(259, 260, True) == [(4, 256)]
    assert fm._input_blocks(512, 512, False) == [(0, 256), (256, 256)]
    assert fm._input_blocks(7, 8, True) is None and fm._input_blocks(100, 100, False) is None


def test_dropin_aliases():
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import scan2cap_b200.dropin as d; d.install()\n"
            "from models.capnet import CapNet\n"
            "from lib.pointnet2.pointnet2_modules import PointnetSAModuleVotes, PointnetFPModule\n"
            "from lib.pointnet2.pytorch_utils import BNMomentumScheduler\n"
            "import pointnet2._ext as _ext\n"
            "assert CapNet.__module__ == 'scan2cap_b200.models.capnet'\n"
            "assert all(hasattr(_ext, n) for n in ['furthest_point_sampling','gather_points','gather_points_grad',"
            "'ball_query','group_points','group_points_grad','three_nn','three_interpolate','three_interpolate_grad'])\n"
            "print('ok')\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "scan2cap_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dirpath, f)


def test_state_dict_contract():
    """Key names of the reference checkpoints."""
    import numpy as np
    from scan2cap_b200 import synthetic
    from scan2cap_b200.data.scannet.model_util_scannet import ScannetDatasetConfig
    from scan2cap_b200.models.capnet import CapNet
    DC = ScannetDatasetConfig()
    vocab, emb, _ = synthetic.make_vocabulary(50)
    m = CapNet(DC.num_class, vocab, emb, DC.num_heading_bin, DC.num_size_cluster, DC.mean_size_arr, input_feature_dim=132,
               num_locals=10, use_topdown=True, query_mode="center", graph_mode="edge_conv", num_graph_steps=2,
               use_relation=True, use_orientation=True)
    sd = m.state_dict()
    for k, shape in {
        "backbone_net.sa1.mlp_module.layer0.conv.weight": (64, 135, 1, 1),
        "backbone_net.sa2.mlp_module.layer0.conv.weight": (128, 131, 1, 1),
        "backbone_net.sa3.mlp_module.layer2.bn.bn.running_mean": (256,),
        "backbone_net.fp1.mlp.layer0.conv.weight": (256, 512, 1, 1),
        "vgen.conv3.weight": (259, 256, 1),
        "proposal.vote_aggregation.mlp_module.layer0.conv.weight": (128, 259, 1, 1),
        "proposal.proposal.6.weight": (97, 128, 1),
        "graph.gc_layers.1.map_edge.2.weight": (128, 128),
        "graph.edge_predict.weight": (7, 128),
        "caption.map_topdown.0.weight": (300, 940),
        "caption.recurrent_cell_2.weight_hh": (1536, 512),
        "caption.classifier.weight": (50, 512),
    }.items():
        assert tuple(sd[k].shape) == shape, k
    detector = [k for k in sd if k.startswith(("backbone_net", "vgen", "proposal"))]
    assert len(detector) == 144  # the reference's PRETRAIN_VOTENET_* checkpoints hold exactly these 144 tensors
