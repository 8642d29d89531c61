"""Compile the C oracle (oracle/pointnet2_oracle.c -> oracle/_build/libs2c_oracle.so).

TEST INFRASTRUCTURE ONLY.  -ffp-contract=off: the oracle spells every fma explicitly so
that its float results are those of the reference's SASS, not of gcc's own contraction.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pointnet2_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libs2c_oracle.so")


def build(force=False, verbose=False):
    if (not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
           "-fno-fast-math", "-mfma", "-o", OUT, SRC, "-lm"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose=True))
