"""In-tree build of the CUDA library (libkmdiff_b200.so) and of the C++ host program for sm_100a.

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
LIB = os.path.join(PKG, "libkmdiff_b200.so")
CLI = os.path.join(PKG, "kmdiff-b200")

# host side: AVX2 (every B200 host has it) for the auto-vectorised sweeps of the eigensolver and the LZ4 / table
# code, but no FMA contraction — host arithmetic must round like the reference's and the oracle's
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xcompiler", "-mavx2", "-Xcompiler", "-ffp-contract=off"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(d: str, exts: tuple[str, ...]) -> list[str]:
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts))


def build_lib(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources(CSRC, (".cu", ".cuh", ".inl", ".hpp")) + [os.path.join(ROOT, "include", "kmdiff_b200.h"), os.path.abspath(__file__)]
    if force or _stale(LIB, srcs):
        cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-o", LIB, os.path.join(CSRC, "kmd_api.cu")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    return LIB


def build_cli(force: bool = False) -> str | None:
    """The C++ host program (`kmdiff-b200 diff ...`), linked against the CUDA library."""
    if not os.path.isdir(HOST):
        return None
    cpps = _sources(HOST, (".cpp",))
    if not cpps:
        return None
    srcs = cpps + _sources(HOST, (".hpp",)) + [os.path.join(ROOT, "include", "kmdiff_b200.h")]
    build_lib(force)
    if force or _stale(CLI, srcs + [LIB]):
        cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", HOST,
               "-o", CLI, *cpps, "-L", PKG, "-lkmdiff_b200", "-Wl,-rpath,$ORIGIN", "-Wl,--export-dynamic",
               "-lpthread", "-ldl"]
        subprocess.check_call(cmd)
    return CLI


def build_all(force: bool = False) -> None:
    build_lib(force)
    build_cli(force)


if __name__ == "__main__":
    import sys
    build_all(force="--force" in sys.argv)
    print(LIB)
