"""Builds libagar_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "agar_b200.cu")
SRC_REPLAY = os.path.join(_HERE, "csrc", "agar_replay.cu")
LIB = os.path.join(_HERE, "libagar_b200.so")
DEPS = [SRC, SRC_REPLAY, os.path.join(_HERE, "csrc", "agar_dev.cuh"), os.path.join(_HERE, "csrc", "agar_bots.cuh"),
        os.path.join(_HERE, "csrc", "agar_simple.cuh")] + [
    os.path.join(os.path.dirname(_HERE), "include", n) for n in ("agar_b200.h", "agar_layout.h", "agar_math.h", "agar_libm_tables.h", "agar_replay.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]


def nvcc_path():
    for p in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if p and os.path.exists(p):
            return p
    raise RuntimeError("nvcc not found")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC, SRC_REPLAY]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
