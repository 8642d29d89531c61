"""The four environment shims that let the reference's UNMODIFIED Python modules
(baseline/_ref/, see install_ref.py) import and run in this image.  None of them touches the reference's
arithmetic; the one substitution of behaviour is labelled below (PyG's MessagePassing.propagate).

  1. ``easydict`` stand-in (not installed); CONF.PATH.SCANNET -> baseline/_ref/data/scannet (the reference
     hard-codes the authors' home directory, lib/config.py:9), CONF.PATH.PRETRAINED -> baseline/_ref/pretrained;
  2. ``pointnet2._ext`` = the reference's own extension compiled for sm_100 (oracle/_ref/pointnet2_ref_ext.so;
     the reference installs it as the namespace package ``pointnet2``, lib/pointnet2/setup.py:21-39);
  3. ``torch_geometric`` is not installable here (no network): a stub provides the names models/graph_module.py
     imports, and **EdgeConv.propagate** (which in the reference re-implements PyG's private propagate with
     version-specific internals, graph_module.py:44-100) is replaced by PyG's documented source_to_target
     semantics: gather x[edge_index[1]] / x[edge_index[0]] -> the reference's own message() -> index_add_ at
     edge_index[1] -> the reference's own update().  This is THE ONLY SUBSTITUTION on the timed path;
  4. modules the path imports but never calls here (trimesh, plyfile, h5py, tensorboardX, matplotlib) are stubbed.

``cpu=True`` additionally neutralises the hard-coded ``.cuda()`` calls (BASELINE config 1: the reference's
capnet_pretrained path on the host cores).

Imports nothing from scan2cap_b200 or from the oracle restatements.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
REF_EXT_SO = os.path.join(ROOT, "oracle", "_ref", "pointnet2_ref_ext.so")
_installed = {}


def load_reference_ext():
    """The reference's CUDA extension (the 9 functions of bindings.cpp:6-19); None if it has not been built."""
    if not os.path.exists(REF_EXT_SO):
        return None
    spec = importlib.util.spec_from_file_location("pointnet2_ref_ext", REF_EXT_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _EasyDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def _from_scipy_sparse_matrix(A):
    A = A.tocoo()
    return torch.from_numpy(np.vstack([A.row, A.col])).long(), torch.from_numpy(A.data)


class _Data(object):
    def __init__(self, x=None, edge_index=None):
        self.x, self.edge_index = x, edge_index


class _MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add"):
        super().__init__()
        self.aggr = aggr


def _propagate(self, edge_index, size=None, **kwargs):
    """PyG flow="source_to_target": x_j = x[edge_index[0]], x_i = x[edge_index[1]], aggregate (add) at edge_index[1]."""
    x = kwargs["x"]
    message = self.message(x_i=x[edge_index[1]], x_j=x[edge_index[0]])
    out = torch.zeros(x.shape[0], message.shape[1], dtype=message.dtype, device=message.device)
    out = out.index_add_(0, edge_index[1], message)
    return self.update(out), message


def install(ext=None, cpu=False, ref_root=REF):
    """Make ``import models.capnet`` etc. resolve to the reference copy under `ref_root`."""
    if _installed:
        return _installed["conf"]
    if not os.path.exists(os.path.join(ref_root, "models", "capnet.py")):
        raise ImportError("baseline/_ref is not installed (run `python -m baseline.install_ref` in the build container)")
    ed = types.ModuleType("easydict")
    ed.EasyDict = _EasyDict
    sys.modules["easydict"] = ed

    if ext is not None:
        pn = types.ModuleType("pointnet2")
        pn._ext = ext
        sys.modules["pointnet2"] = pn
        sys.modules["pointnet2._ext"] = ext
    else:
        import builtins
        builtins.__POINTNET2_SETUP__ = True  # the reference's own escape hatch (pointnet2_utils.py:28): no native ops

    tg = types.ModuleType("torch_geometric")
    sub = {n: types.ModuleType("torch_geometric." + n) for n in ("utils", "data", "nn", "typing")}
    sub["utils"].from_scipy_sparse_matrix = _from_scipy_sparse_matrix
    sub["utils"].add_self_loops = sub["utils"].degree = None
    sub["data"].Data, sub["data"].DataLoader = _Data, None
    sub["nn"].MessagePassing, sub["nn"].GCNConv = _MessagePassing, None
    sub["typing"].Adj = sub["typing"].Size = None
    sys.modules["torch_geometric"] = tg
    for n, m in sub.items():
        setattr(tg, n, m)
        sys.modules[m.__name__] = m

    for missing in ("trimesh", "plyfile", "h5py", "tensorboardX", "matplotlib", "matplotlib.pyplot"):
        try:
            __import__(missing)
        except Exception:
            stub = types.ModuleType(missing)
            stub.PlyData = stub.PlyElement = stub.SummaryWriter = None
            stub.cm = types.SimpleNamespace(jet=None)
            stub.pyplot = stub
            sys.modules[missing] = stub

    if cpu:
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.cuda.FloatTensor = torch.FloatTensor

    sys.path.insert(0, ref_root)
    sys.path.insert(0, os.path.join(ref_root, "lib", "pointnet2"))  # `import pointnet2_utils` (pointnet2_modules.py:21)
    import lib.config as cfg
    cfg.CONF.PATH.BASE = ref_root
    cfg.CONF.PATH.SCANNET = os.path.join(ref_root, "data", "scannet")
    cfg.CONF.PATH.SCANNET_META = os.path.join(ref_root, "data", "scannet", "meta_data")
    cfg.CONF.PATH.PRETRAINED = os.path.join(ref_root, "pretrained")

    import models.graph_module as gm
    gm.EdgeConv.propagate = _propagate
    _installed["conf"] = cfg.CONF
    return cfg.CONF
