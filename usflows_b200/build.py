"""Builds libusflows_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m usflows_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so lands in usflows_b200/lib/ (git-ignored, shipped to the GPU
box with the repo snapshot).  No JIT cache, no pip install.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libusflows_b200.so")
STAMP = os.path.join(LIBDIR, "build.stamp")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _sources():
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))
    files.append(os.path.join(INCLUDE, "usflows_b200.h"))
    return files


def _digest() -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in _sources():
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode() + b"\0" + fh.read())   # path independent: the .so built here is valid on the GPU box
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if sources changed; returns the library path."""
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(LIBDIR, exist_ok=True)
    units = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    objs = [os.path.join(LIBDIR, u[:-3] + ".o") for u in units]

    def compile_one(pair):
        src, obj = pair
        cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-c", "-o", obj, os.path.join(CSRC, src)]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        return " ".join(cmd) + "\n" + proc.stdout + proc.stderr, proc.returncode

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(units)) as pool:      # translation units compile in parallel
        results = list(pool.map(compile_one, zip(units, objs)))
    log = "\n".join(r[0] for r in results)
    rc = max(r[1] for r in results)
    if rc == 0:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "--cudart", "shared", "-o", LIB] + objs
        proc = subprocess.run(cmd, capture_output=True, text=True)
        log += "\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr
        rc = proc.returncode
    with open(os.path.join(LIBDIR, "build.log"), "w") as f:
        f.write(log)
    if rc != 0:
        raise RuntimeError("nvcc failed:\n" + log[-8000:])
    if verbose:
        print(log)
    with open(STAMP, "w") as f:
        f.write(_digest())
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
