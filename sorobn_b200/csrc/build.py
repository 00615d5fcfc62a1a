"""Build libsorobn_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

No torch, no pybind: the library's only dependency is the CUDA runtime (linked
statically), so it loads with ctypes from any process.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, "libsorobn_b200.so")
SOURCES = [os.path.join(HERE, "sbn_api.cu")]
HEADERS = [os.path.join(HERE, "sbn_kernels.cuh"), os.path.join(HERE, "sbn_gibbs.cuh"), os.path.join(os.path.dirname(PKG), "include", "sorobn_b200.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(f) <= t for f in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    cmd = [
        nvcc_path(), "-shared", "-Xcompiler", "-fPIC", "-O3", "-std=c++17", "-lineinfo",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-Xptxas", "-v" if verbose else "-O3",
        "-o", LIB, *SOURCES,
    ]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libsorobn_b200.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
