"""Build libs2c.so (hand-written sm_100a CUDA kernels + the C ABI of include/s2c.h) IN-TREE.

    python -m scan2cap_b200.build [-f] [-v]

nvcc cross-compiles for sm_100a without a GPU.  The library lands next to this file
(scan2cap_b200/libs2c.so): git-ignored, but shipped to the GPU box by gpurun.
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libs2c.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
          "--expt-relaxed-constexpr", "-Xcompiler", "-fvisibility=hidden"]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "s2c.h"), __file__]


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    newest = max(os.path.getmtime(p) for p in [src] + _deps())
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return obj, ""
    cmd = [NVCC] + ARCH + CFLAGS + ["-c", src, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, p.stdout, p.stderr))
    return obj, p.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    objs = [r[0] for r in res]
    log = "".join(r[1] for r in res)
    if log:
        with open(os.path.join(OBJ, "ptxas.log"), "a") as f:
            f.write(log)
        if verbose:
            print(log)
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
