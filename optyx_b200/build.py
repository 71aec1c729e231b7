"""Build the CUDA shared library in-tree (``optyx_b200/lib/liboptyx_b200.so``).

nvcc cross-compiles sm_100a without a GPU.  The library is git-ignored but
travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "liboptyx_b200.so")
STAMP = os.path.join(LIBDIR, "liboptyx_b200.stamp")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

SOURCES = ["abi.cu", "leaf_build.cu", "contract_direct.cu", "contract_gemm.cu", "contract_gemm_c64.cu", "contract_ctile.cu", "contract_fma.cu", "contract_small.cu", "permanent.cu",
           "probs.cu", "compact.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
    "-std=c++17", "--shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _digest() -> str:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    files.append(os.path.join(INCLUDE, "optyx_b200.h"))
    for path in files:
        with open(path, "rb") as fh:
            h.update(path.encode())
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a into one shared library; returns its path."""
    os.makedirs(LIBDIR, exist_ok=True)
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == digest:
                return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-o", LIB,
           *[os.path.join(CSRC, s) for s in SOURCES]]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(LIBDIR, "build.log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if verbose or proc.returncode != 0:
        sys.stderr.write(log)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building liboptyx_b200.so (see above)")
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
