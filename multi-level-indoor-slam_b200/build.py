"""Build libsemgate.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python multi-level-indoor-slam_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "semgate", "libsemgate.so")
SOURCES = ["api.cu", "gated_topk.cu", "kernels.cu", "spatial.cu", "rerank.cu", "stream_query.cu"]
HEADERS = ["common.cuh", "ptx.cuh", "gated_topk.cuh", "merge.cuh", "launch.h", os.path.join("..", "..", "include", "semgate.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; libsemgate cannot be built")


def needs_build() -> bool:
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    cmd = [
        _nvcc(), "-shared", "-std=c++17", "-O3", "-lineinfo",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-Xcompiler", "-fPIC,-O3,-Wall",
        "-Xptxas", "-v" if verbose else "-O3",
        "-DSEMGATE_BUILD",
        "-o", OUT,
    ] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libsemgate.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
