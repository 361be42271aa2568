"""Builds conzic_b200/libconzic.so (hand-written sm_100a CUDA behind the C ABI of include/conzic.h)
and, for the test harness only, oracle/_cabi checks.  nvcc cross-compiles without a GPU.

    python -m conzic_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libconzic.so")
STAMP = os.path.join(PKG, ".libconzic.stamp")
OBJDIR = os.path.join(PKG, "_obj")
SOURCES = ["engine.cu", "gemm.cu", "transformer_ops.cu", "select_ops.cu", "cert_ops.cu", "text_ops.cu", "image_ops.cu"]
HEADERS = ["kernels.h", "ptx.cuh", "select_common.cuh", "text_pipeline.cuh", os.path.join(ROOT, "include", "conzic.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libconzic.so cannot be built (there is no CPU fallback)")


def _digest() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        p = f if os.path.isabs(f) else os.path.join(CSRC, f)
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    nvcc = find_nvcc()
    os.makedirs(OBJDIR, exist_ok=True)

    def compile_one(src):  # one translation unit per nvcc process, all of them at once
        obj = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, src)]
        return obj, subprocess.run(cmd, capture_output=True, text=True)

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    for _, res in results:
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
    if any(res.returncode != 0 for _, res in results):
        raise RuntimeError("nvcc failed building libconzic.so")
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] +
                         [obj for obj, _ in results], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libconzic.so")
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
