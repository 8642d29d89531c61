"""Build the UNMODIFIED reference CUDA extension (lib/pointnet2/_ext_src) for sm_100.

TEST INFRASTRUCTURE ONLY.  The sources are compiled where they lie under
/root/reference (never copied into this repo); only build outputs land in
oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).  The result,
oracle/_ref/pointnet2_ref_ext.so, is the reference's own kernels
(sampling_gpu.cu, ball_query_gpu.cu, group_points_gpu.cu, interpolate_gpu.cu)
and is used by tests/ and bench.py --impl reference as the bit-exact GPU oracle.

We do not run the reference's setup.py (its arch list "3.7+PTX;...;7.5",
lib/pointnet2/setup.py:17, no longer exists in CUDA 12.9); we call
torch.utils.cpp_extension.load on the same 9 source files with the same -O3 flags
(setup.py:30-33).
"""
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/lib/pointnet2/_ext_src"
OUT = os.path.join(HERE, "_ref")
NAME = "pointnet2_ref_ext"


def build(verbose=False):
    so = os.path.join(OUT, NAME + ".so")
    if not os.path.isdir(REF_SRC):
        return so if os.path.exists(so) else None
    srcs = sorted(glob.glob(os.path.join(REF_SRC, "src", "*.cpp")) +
                  glob.glob(os.path.join(REF_SRC, "src", "*.cu")))
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
        return so
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0"
    from torch.utils.cpp_extension import load
    load(name=NAME, sources=srcs,
         extra_include_paths=[os.path.join(REF_SRC, "include")],
         extra_cflags=["-O3"], extra_cuda_cflags=["-O3"],
         build_directory=OUT, verbose=verbose, is_python_module=False)
    return so


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
