"""Install the UNMODIFIED reference Python path + the shipped VoteNet checkpoints under baseline/_ref/.

    python -m baseline.install_ref

baseline/_ref/ is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so the copy
travels to the GPU box, where /root/reference does not exist.  The reference has no setup.py for its Python
layers (only lib/pointnet2/setup.py, whose arch list no longer exists in CUDA 12.9 -- see oracle/build_ref.py),
so "install" = a verbatim file copy of exactly the modules the hot path imports:

    lib/*.py, lib/pointnet2/*.py, models/*.py, utils/*.py, data/scannet/*.py, data/scannet/meta_data/*,
    pretrained/PRETRAIN_VOTENET_XYZ{,_MULTIVIEW_NORMAL}/model.pth

A no-op (returns the existing copy, or None) when /root/reference is absent.
"""
import filecmp
import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
PATTERNS = ["lib/*.py", "lib/pointnet2/*.py", "models/*.py", "utils/*.py", "data/scannet/*.py",
            "data/scannet/meta_data/*", "pretrained/PRETRAIN_VOTENET_XYZ/model.pth",
            "pretrained/PRETRAIN_VOTENET_XYZ_MULTIVIEW_NORMAL/model.pth"]


def installed():
    return os.path.exists(os.path.join(OUT, "models", "capnet.py"))


def install(verbose=False):
    if not os.path.isdir(REF):
        return OUT if installed() else None
    n = 0
    for pat in PATTERNS:
        for src in sorted(glob.glob(os.path.join(REF, pat))):
            if not os.path.isfile(src):
                continue
            dst = os.path.join(OUT, os.path.relpath(src, REF))
            if os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False):
                continue
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            os.chmod(dst, 0o644)
            n += 1
    if verbose:
        print("baseline/_ref: %d file(s) copied from %s" % (n, REF))
    return OUT


if __name__ == "__main__":
    print(install(verbose=True))
    sys.exit(0 if installed() else 1)
