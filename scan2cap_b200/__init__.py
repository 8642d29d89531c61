"""scan2cap_b200 -- Blackwell-native (sm_100a) implementation of the Scan2Cap point-cloud -> caption hot path.

Layout (mirrors the reference's own module paths so that callers switch by import only):
    scan2cap_b200/csrc/             hand-written CUDA kernels + the C ABI of include/s2c.h  -> libs2c.so
    scan2cap_b200/_lib.py           ctypes binding of libs2c.so (fails loudly if the library is missing)
    scan2cap_b200/lib/pointnet2/    _ext / pointnet2_utils / pointnet2_modules / pytorch_utils mirrors
    scan2cap_b200/models/           backbone / voting / proposal / graph / caption / CapNet mirrors
    scan2cap_b200/dropin.py         installs the mirrors under the reference's import names
"""
__version__ = "0.1.0"
