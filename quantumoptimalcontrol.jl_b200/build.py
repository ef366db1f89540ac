"""Builds csrc/ into lib/libqoc_b200.so for sm_100a with nvcc (in-tree, so the .so travels with the repo
snapshot to the GPU box).  nvcc cross-compiles without a GPU."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libqoc_b200.so")
SOURCES = [os.path.join(CSRC, "qoc_api.cu")]
# every header of csrc/ (globbed: a header missing from a hand-kept list would let an edited kernel run against a stale binary)
HEADERS = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [
    os.path.join(os.path.dirname(HERE), "include", "qoc_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "--cudart", "static"]


STAMP = LIB + ".srchash"


def source_hash() -> str:
    """sha256 over the sources and the flags: content, not mtimes (the gpurun snapshot does not promise to keep mtimes, and N
    ranks of one torchrun must agree on whether the shipped binary is current)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in SOURCES + sorted(HEADERS):
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    tmp = LIB + f".tmp{os.getpid()}"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    os.replace(tmp, LIB)   # atomic: a concurrent loader sees the old or the new library, never a torn one
    with open(STAMP + f".tmp{os.getpid()}", "w") as fh:
        fh.write(source_hash())
    os.replace(STAMP + f".tmp{os.getpid()}", STAMP)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
