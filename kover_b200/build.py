"""Builds libkoverb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "kover_b200.cu")
HDR = os.path.join(ROOT, "include", "kover_b200.h")
INC = os.path.join(HERE, "csrc", "kv_inflate.cuh")
LIB = os.path.join(HERE, "libkoverb200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "--fmad=false",                      # the utility / Gini arithmetic must never be contracted
    "-Xcompiler", "-fPIC,-ffp-contract=off,-pthread",
    "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.isfile(cand)):
            return cand
    return "nvcc"


def is_stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in (SRC, HDR, INC))


def build_native(force=False, verbose=False):
    """Compile the CUDA library if missing or older than its sources.  Returns the .so path."""
    if not force and not is_stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), SRC, "-o", LIB]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
