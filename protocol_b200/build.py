"""Builds libprime_match.so (CUDA engine + host helpers) in-tree for sm_100a.

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the
GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libprime_match.so")
SOURCES = [os.path.join(HERE, "csrc", "pm_engine.cu"), os.path.join(HERE, "csrc", "pm_host.cpp"),
           os.path.join(HERE, "csrc", "pm_plugin.cpp")]
# every header under csrc/ and include/ is a dependency: a stale .so on the GPU box would silently run old kernels
DEPS = SOURCES + sorted(
    os.path.join(d, f)
    for d in (os.path.join(HERE, "csrc"), os.path.join(ROOT, "include"))
    for f in os.listdir(d) if f.endswith((".cuh", ".hpp", ".h"))
)

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-shared",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True, cwd=HERE)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
