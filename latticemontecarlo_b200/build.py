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
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))) + [os.path.join("..", "..", "include", "lmc_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
         "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "-ccbin", "/usr/bin/g++", "--expt-relaxed-constexpr"]


CLI_SRC = os.path.join(CSRC, "host", "lmc_cli.cpp")
CLI_OUT = os.path.join(HERE, "lmc_b200.exe")


def build_cli(force: bool = False) -> str:
    """`lmc_b200.exe -p <param file>`: the reference's command-line surface (host C++ over the C ABI)."""
    if not force and os.path.exists(CLI_OUT) and os.path.getmtime(CLI_OUT) > max(os.path.getmtime(CLI_SRC), os.path.getmtime(OUT)):
        return CLI_OUT
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", CLI_SRC, "-L" + HERE, "-llmc_b200", "-lz", "-Wl,-rpath," + HERE, "-o", CLI_OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building lmc_b200.exe")
    return CLI_OUT


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        build_cli()
        return OUT
    extra = os.environ.get("LMC_NVCC_EXTRA", "").split()          # e.g. -DLMC_CMC_PROFILE for the clock64 phase profile
    cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building liblmc_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    build_cli(force=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
