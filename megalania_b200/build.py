"""Builds the product artefacts in-tree (they travel to the GPU box with the snapshot):

  megalania_b200/_build/libmegalania_cuda.so   CUDA kernels + C ABI, sm_100a only
  megalania_b200/_build/megalania              C host CLI (drop-in for the reference binary)

    python -m megalania_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
OUT = os.path.join(PKG, "_build")
LIB = os.path.join(OUT, "libmegalania_cuda.so")
CLI = os.path.join(OUT, "megalania")

CUDA_SOURCES = [os.path.join(PKG, "csrc", "mg_api.cu")]
CUDA_DEPS = [os.path.join(PKG, "csrc", f) for f in ("mg_device.cuh", "mg_finder.cuh", "mg_kernels.cuh", "mg_comm.inc")] + [
    os.path.join(ROOT, "include", f) for f in ("megalania_cuda.h", "output_interface.h", "encoder_interface.h")]
HOST_SOURCES = [os.path.join(PKG, "host", f) for f in ("main.c", "host_io.c")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-ldl"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    if force or _stale(LIB, CUDA_SOURCES + CUDA_DEPS + [os.path.abspath(__file__)]):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + CUDA_SOURCES
        subprocess.run(cmd, check=True)
    return LIB


def build_cli(force: bool = False) -> str | None:
    srcs = [s for s in HOST_SOURCES if os.path.exists(s)]
    if len(srcs) != len(HOST_SOURCES):
        return None
    if force or _stale(CLI, srcs + [LIB]):
        cmd = ["gcc", "-O2", "-std=gnu11", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), "-o", CLI] + srcs + [
            "-L", OUT, "-lmegalania_cuda", "-Wl,-rpath,$ORIGIN", "-lm", "-lpthread"]
        subprocess.run(cmd, check=True)
    return CLI


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_library(force, verbose)
    build_cli(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
