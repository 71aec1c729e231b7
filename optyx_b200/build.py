"""Build the CUDA shared library in-tree (``optyx_b200/lib/liboptyx_b200.so``).

nvcc cross-compiles sm_100a without a GPU.  The library is git-ignored but
travels to the GPU box with the repo snapshot.  Every ``.cu`` is compiled to its
own object (in parallel, cached by a digest of the source, the shared headers and
the flags) and the objects are linked into one shared library; the library's
stamp is the digest of ALL sources, keyed by repo-relative names so that it still
matches after the tree has been copied to another machine.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "liboptyx_b200.so")
STAMP = os.path.join(LIBDIR, "liboptyx_b200.stamp")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

SOURCES = ["abi.cu", "leaf_build.cu", "contract_direct.cu", "contract_gemm.cu", "contract_gemm_ws.cu", "contract_gemm_c64.cu",
           "contract_ctile.cu", "contract_fma.cu", "contract_small.cu", "permanent.cu", "probs.cu",
           "compact.cu", "plan_runner.cu", "planner_native.cu", "compress.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
    "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _headers():
    hs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    return hs + [os.path.join(INCLUDE, "optyx_b200.h")]


def _hash_files(paths) -> str:
    h = hashlib.sha256()
    for path in paths:
        with open(path, "rb") as fh:
            h.update(os.path.basename(path).encode())      # relative name, not the absolute path
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _digest() -> str:
    """Digest of every source, header and flag that goes into the library."""
    return _hash_files([os.path.join(CSRC, s) for s in _sources()] + _headers())


def stamp_matches() -> bool:
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == _digest()


def _compile_one(nvcc: str, src: str):
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJDIR, src[:-3] + ".o")
    tag = obj + ".stamp"
    digest = _hash_files([path] + _headers())
    if os.path.exists(obj) and os.path.exists(tag):
        with open(tag) as fh:
            if fh.read().strip() == digest:
                return obj, "", 0
    cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", "-o", obj, path]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode == 0:
        with open(tag, "w") as fh:
            fh.write(digest)
    return obj, " ".join(cmd) + "\n" + proc.stdout + proc.stderr, proc.returncode


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a into one shared library; returns its path."""
    os.makedirs(OBJDIR, exist_ok=True)
    if not force and stamp_matches():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if force:
        for f in os.listdir(OBJDIR):
            os.remove(os.path.join(OBJDIR, f))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(lambda s: _compile_one(nvcc, s), _sources()))
    log = "".join(r[1] for r in results)
    failed = any(r[2] != 0 for r in results)
    if not failed:
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", LIB,
               *[r[0] for r in results]]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + proc.stdout + proc.stderr
        failed = proc.returncode != 0
    with open(os.path.join(LIBDIR, "build.log"), "w") as fh:
        fh.write(log)
    if verbose or failed:
        sys.stderr.write(log)
    if failed:
        raise RuntimeError("nvcc failed building liboptyx_b200.so (see above)")
    with open(STAMP, "w") as fh:
        fh.write(_digest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
