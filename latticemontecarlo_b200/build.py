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
SOURCES = ["engine.cu", "cmc_domain.cu", "grouping.cu", "tables.cpp"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))) + [os.path.join("..", "..", "include", "lmc_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "-ccbin", "/usr/bin/g++", "--expt-relaxed-constexpr"]
OBJ_DIR = os.path.join(HERE, "build")


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
    """One object per source (compiled in parallel, rebuilt only when the source or a header is newer), then one link."""
    if not force and not needs_build():
        build_cli()
        return OUT
    from concurrent.futures import ThreadPoolExecutor
    extra = os.environ.get("LMC_NVCC_EXTRA", "").split()          # e.g. -DLMC_CMC_PROFILE for the clock64 phase profile
    os.makedirs(OBJ_DIR, exist_ok=True)
    newest_header = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    stamp = os.path.join(OBJ_DIR, "flags.txt")
    flags_text = " ".join(FLAGS + extra)
    same_flags = os.path.exists(stamp) and open(stamp).read() == flags_text

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
        path = os.path.join(CSRC, src)
        if not force and same_flags and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), newest_header):
            return obj, ""
        cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed compiling " + src)
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    with open(stamp, "w") as f:
        f.write(flags_text)
    if verbose:
        for _, err in results:
            sys.stderr.write(err)
    res = subprocess.run([NVCC, "--shared", "-ccbin", "/usr/bin/g++"] + [o for o, _ in results] + ["-o", OUT], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking liblmc_b200.so")
    build_cli(force=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
