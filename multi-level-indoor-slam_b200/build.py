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
HEADERS = ["common.cuh", "ptx.cuh", "gated_topk.cuh", "merge.cuh", "sortnet.cuh", "launch.h", os.path.join("..", "..", "include", "semgate.h")]


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


TORCH_OPS_SRC = os.path.join(CSRC, "torch_ops.cpp")
TORCH_OPS_OUT = os.path.join(HERE, "semgate", "libsemgate_torch.so")


def build_torch_ops(force: bool = False) -> str:
    """torch.ops.semgate.*: a thin C++ op library over libsemgate.so's C ABI (no kernels of its own),
    compiled in-tree against the running torch's headers."""
    build(force=False)
    deps = [TORCH_OPS_SRC, os.path.join(CSRC, "..", "..", "include", "semgate.h"), os.path.abspath(__file__)]
    if not force and os.path.isfile(TORCH_OPS_OUT) and all(os.path.getmtime(d) <= os.path.getmtime(TORCH_OPS_OUT) for d in deps):
        return TORCH_OPS_OUT
    import torch
    from torch.utils import cpp_extension
    cuda_home = os.path.dirname(os.path.dirname(_nvcc()))
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-shared", "-fPIC", "-std=c++17", "-O2", "-Wall", "-Wno-unknown-pragmas",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", TORCH_OPS_SRC]
    cmd += ["-I" + p for p in cpp_extension.include_paths()] + ["-I" + os.path.join(cuda_home, "include")]
    cmd += ["-L" + torch_lib, "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_cuda", "-lc10_cuda",
            "-L" + os.path.dirname(OUT), "-lsemgate", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + torch_lib, "-o", TORCH_OPS_OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building libsemgate_torch.so")
    return TORCH_OPS_OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    if "--torch-ops" in sys.argv:
        print(build_torch_ops(force="--force" in sys.argv))
