"""TEST INFRASTRUCTURE ONLY: g++ build of the CUDA sources on the SIMT emulator.

    python tests/emu/build_emu.py [name] [extra flags, e.g. -DDZ_STEP_TILED=1]

writes tests/emu/_build/libdantzig_b200_emu[_name].so (git-ignored).  The library
exports the C ABI of include/dantzig_b200.h but runs every "kernel" on the CPU
through tests/emu/simt_emu.h; tests load it with the DZ_LIB override to check the
kernel logic without a GPU.  Nothing under dantzig_b200/ ever builds or loads it.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "dantzig_b200", "csrc")


def build(name: str = "", flags=(), force: bool = False) -> str:
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libdantzig_b200_emu%s.so" % ("_" + name if name else ""))
    srcs = [os.path.join(CSRC, f) for f in ("dz_kernel.cu", "dz_core.cu", "dz_grid.cu", "dz_fast.cu", "dz_capi.cu", "dz_lower.cpp")]
    srcs.append(os.path.join(HERE, "simt_emu.cpp"))
    deps = srcs + [os.path.join(HERE, "simt_emu.h"), os.path.join(CSRC, "dz_internal.h"), os.path.join(CSRC, "dz_device.cuh"),
                   os.path.join(ROOT, "include", "dantzig_b200.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    cmd = [os.environ.get("CXX", "g++"), "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
           "-DDZ_EMU", "-I", HERE, *flags, "-x", "c++", *srcs, "-o", out]
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    print(build(sys.argv[1] if len(sys.argv) > 1 else "", sys.argv[2:], force=True))
