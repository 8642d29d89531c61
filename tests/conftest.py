import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_reference_ext():
    """The UNMODIFIED reference CUDA extension built by oracle/build_ref.py (None if not built)."""
    so = os.path.join(ROOT, "oracle", "_ref", "pointnet2_ref_ext.so")
    if not os.path.exists(so):
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location("pointnet2_ref_ext", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def ref_ext():
    return load_reference_ext()


@pytest.fixture(scope="session")
def ext():
    from scan2cap_b200.lib.pointnet2 import _ext
    return _ext


@pytest.fixture(scope="session")
def oracle():
    from oracle import native
    return native
