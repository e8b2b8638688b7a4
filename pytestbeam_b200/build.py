"""Build ``libptb_b200.so`` in-tree with nvcc for sm_100a (``python -m pytestbeam_b200.build``)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "ptb_kernels.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "ptb_device.cuh"), os.path.join(ROOT, "include", "ptb.h")]
OUT_DIR = os.path.join(HERE, "_lib")
OUT = os.path.join(OUT_DIR, "libptb_b200.so")

# -fmad=false: the exact build keeps the reference's IEEE operation order (no FMA contraction);
# -lineinfo so that ncu's source page maps back to the .cu files.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(exe):
        raise RuntimeError("nvcc not found")
    return exe


def up_to_date() -> bool:
    return os.path.isfile(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [nvcc()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-o", OUT, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
