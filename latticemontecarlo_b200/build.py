"""Build liblmc_b200.so (hand-written sm_100a CUDA + C++ host) in-tree with nvcc.

    python -m latticemontecarlo_b200.build

The shared library is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "liblmc_b200.so")
SOURCES = ["engine.cu", "tables.cpp"]
HEADERS = ["engine.h", "kernels.cuh", "kmc_kernels.cuh", "cmc_kernels.cuh", "tables.h", "lattice.h", "device_tables.h", os.path.join("..", "..", "include", "lmc_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
         "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "-ccbin", "/usr/bin/g++", "--expt-relaxed-constexpr"]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building liblmc_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
