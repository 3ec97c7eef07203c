"""Build library variants for tools/gpu_ab.py (run here, on the CPU: nvcc cross-compiles).

    python tools/build_variants.py base bsub=-DDZ_BSUB_COMPACT=1 "both=-DDZ_BSUB_COMPACT=1 -DDZ_BSUB_U32=1"
    gpurun -- 'AB_FAST=1 python tools/gpu_ab.py build/ab/bsub.so build/ab/both.so build/ab/base.so'

`name` alone builds the shipped configuration; `name=flags` adds nvcc flags (the
DZ_* switches listed at the top of dantzig_b200/csrc/dz_kernel.cu).  Output goes to
build/ab/ (git-ignored, travels with gpurun).  Prints registers and spills of the
warp-per-LP kernel for each variant.
"""
import os, re, subprocess, sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dantzig_b200 import build  # noqa: E402


def one(spec):
    name, _, flags = spec.partition("=")
    out = os.path.join(ROOT, "build", "ab", name + ".so")
    srcs = [os.path.join(build.CSRC, s) for s in build.LIB_SOURCES]
    cmd = [build._nvcc(), *build.NVCC_FLAGS, "-Xptxas", "-v", *flags.split(), "-shared", "-o", out, *srcs]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=build.PKG)
    if r.returncode:
        return name, "FAILED\n" + r.stderr[-2000:]
    m = re.search(r"ILi1ELb1ELi4E.*?\n\s*(\d+ bytes stack frame, \d+ bytes spill stores, \d+ bytes spill loads)\n.*?Used (\d+) registers",
                  r.stderr, re.S)
    return name, ("%s registers, %s" % (m.group(2), m.group(1))) if m else "built"


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "build", "ab"), exist_ok=True)
    with ThreadPoolExecutor(4) as ex:
        for name, msg in ex.map(one, sys.argv[1:]):
            print("%-12s %s" % (name, msg))
