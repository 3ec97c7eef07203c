"""In-tree build of the native pieces (no JIT cache: the .so files travel with
the repo snapshot to the GPU box).

    libdantzig_b200.so   CUDA kernels + C ABI (nvcc, sm_100a, --fmad=false)
    rust.<abi>.so        the drop-in ``dantzig.rust`` extension module (g++/pybind11)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libdantzig_b200.so")
EXT = os.path.join(PKG, "rust" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # the exact path may never contract a*b-c into an FMA
    "--fmad=false",
    "-Xcompiler", "-fPIC",
]
LIB_SOURCES = ["dz_kernel.cu", "dz_core.cu", "dz_grid.cu", "dz_fast.cu", "dz_capi.cu", "dz_lower.cpp"]
LIB_HEADERS = ["dz_internal.h", "dz_device.cuh", os.path.join("..", "..", "include", "dantzig_b200.h")]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in LIB_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in LIB_HEADERS]
    if force or _stale(LIB, deps):
        cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-o", LIB, *srcs]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True, cwd=PKG)
    return LIB


def build_extension(force: bool = False) -> str:
    """The ``dantzig.rust`` replacement module (host C++ above the C ABI)."""
    import pybind11

    src = os.path.join(CSRC, "dz_rustmod.cpp")
    if force or _stale(EXT, [src, LIB, os.path.join(PKG, "..", "include", "dantzig_b200.h")]):
        cmd = [
            os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared",
            "-fvisibility=hidden",
            "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"],
            src, "-o", EXT, "-L", PKG, "-ldantzig_b200", "-Wl,-rpath,$ORIGIN",
        ]
        subprocess.run(cmd, check=True, cwd=PKG)
    return EXT


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_library(force=force, verbose=verbose)
    build_extension(force=force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
    print(EXT)
