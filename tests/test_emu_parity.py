"""The CUDA kernel source, compiled by g++ on the SIMT emulator of tests/emu and
checked against the golden fixtures on the CPU.

This is test infrastructure: tests/emu/simt_emu.h runs every CUDA thread as a
fiber and resolves warp collectives and block barriers, so the SAME kernel code
(dantzig_b200/csrc/dz_kernel.cu, built with -DDZ_EMU) executes here without a
GPU.  It checks the kernel's logic in every launch shape, the on-chip core kernel
(dz_core.cu) included, and the build switches before they ever see a GPU.
The product never builds or loads the emulated library; the parity tests proper
are the `-m gpu` ones.
"""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
ALL_SHAPES = "0:0,-1:0,1:2,3:1,2:3,0:4"  # the six launch shapes of tests/test_gpu_parity.py


def _build(name="", flags=()):
    spec = importlib.util.spec_from_file_location("build_emu", os.path.join(EMU, "build_emu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(name, list(flags))


def _run(lib, *args):
    env = dict(os.environ, DZ_LIB=lib, DZ_LIB_TEST_ONLY="1")
    r = subprocess.run([sys.executable, os.path.join(EMU, "run_child.py"), *args], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("EMU ")][-1]
    return json.loads(line[4:])


def _assert_clean(res):
    bad = {k: v for k, v in res.items() if v[0] != 0}
    assert not bad, bad
    assert res and all(v[1] > 0 for v in res.values())


def test_shipped_kernel_every_launch_shape():
    lib = _build()
    _assert_clean(_run(lib, "golden", ALL_SHAPES, "tiny_4x6:8", "small_8x16:4", "mixed_9x12:6",
                       "mixed_20x40:2", "packing_24x48:1"))


def test_shipped_kernel_warp_fast_paths():
    """ceil(m_int/32) = 1..4 (64-register kernel) and 6 (128-register kernel), one warp per LP."""
    lib = _build()
    _assert_clean(_run(lib, "golden", "-1:0", "c2_32x64:2", "small_40x80:1", "c2_false_unbounded:1"))


def test_reference_kats_through_the_c_abi():
    lib = _build()
    _assert_clean(_run(lib, "kats"))


def test_core_kernel_shapes_and_hand_over():
    """The on-chip coupled-core kernel (basis_home=4): every rows-per-lane class it is
    instantiated for that a CPU can afford, the false-unbounded LPs, and a shared-memory
    budget so small that core rows overflow into the HBM workspace (ctas_per_sm=16)."""
    lib = _build()
    _assert_clean(_run(lib, "golden", "0:4", "c2_32x64:2", "small_40x80:1", "c2_false_unbounded:2",
                       "packing_24x48:2"))
    _assert_clean(_run(lib, "golden", "0:4:16", "mixed_20x40:2", "c2_32x64:1"))


def test_grid_kernel_on_small_grids():
    """The whole-GPU single-LP kernel (basis_home=5) as a cooperative grid of three CTAs of two
    warps (the emulator schedules all CTAs' fibers and resolves the grid barrier): master/worker
    job loop, look-ahead windows, jump bookkeeping, level-scheduled back-substitution with spin
    waits.  Then the same with a 3-entry heap, 32-column update tiles and a 40-column window, so
    that the heap-overflow fallback, multi-tile updates and window refreshes all run."""
    lib = _build()
    _assert_clean(_run(lib, "golden", "2:5:3", "tiny_4x6:8", "mixed_9x12:6", "mixed_20x40:2", "c2_32x64:1",
                       "c2_false_unbounded:1", "packing_24x48:1"))
    small = _build("gridsmall", ["-DDZ_GRID_HEAP_CAP=3", "-DDZ_GRID_TILE=32", "-DDZ_GRID_WIN=40"])
    _assert_clean(_run(small, "golden", "2:5:3", "tiny_4x6:8", "mixed_9x12:6", "mixed_20x40:1"))
    _assert_clean(_run(small, "golden", "4:5:2", "mixed_9x12:4", "c2_32x64:1"))


def test_grid_and_core_kernels_on_integer_transportation_lps():
    """Ties in every pivot search (integer data): the look-ahead windows must break them by the
    rows' CURRENT positions.  Full solves against the sparse oracle."""
    lib = _build()
    env = dict(os.environ, DZ_LIB=lib, DZ_LIB_TEST_ONLY="1")
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from dantzig_b200 import generate, Template, solve_batch\n"
        "from oracle import dzo_py\n"
        "bad = 0\n"
        "for seed, shape in ((0, (10, 10, 40, 1)), (1, (20, 20, 120, 3))):\n"
        "    model = generate.transportation_model(seed, *shape)\n"
        "    t = Template(model); th = t.pack_theta(model)[None, :]\n"
        "    o = dzo_py.lower(model).solve(dzo_py.SPARSE)\n"
        "    for kw in (dict(worker_warps=2, basis_home=5, ctas_per_sm=3), dict(basis_home=4)):\n"
        "        r = solve_batch(t, th, **kw)\n"
        "        bad += (int(r.status[0]), int(r.pivots[0]), int(r.trace_hash[0])) != (o.status, o.pivots, o.trace_hash)\n"
        "        bad += not np.array_equal(r.x_basic[0], o.x_basic)\n"
        "print('BAD', bad)\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "BAD 0" in r.stdout


@pytest.mark.parametrize("name,flags", [
    ("allinone", ["-DDZ_KERNEL_PER_NR=0"]),
    ("prof", ["-DDZ_STEP_PROFILE=1"]),
])
def test_build_variants_keep_parity(name, flags):
    """The remaining build switches (tools/README.md) on the shapes they touch."""
    lib = _build(name, flags)
    _assert_clean(_run(lib, "golden", "-1:0,0:0", "tiny_4x6:8", "mixed_9x12:6", "mixed_20x40:1"))
    _assert_clean(_run(lib, "golden", "-1:0", "packing_24x48:1", "c2_32x64:1", "small_40x80:1"))


def test_gpu_suite_subset_on_the_emulator():
    """The ctypes-facing part of tests/test_gpu_parity.py (KATs, ragged and random models,
    integer-data batches in three launch shapes, the transportation family, pivot cap,
    array-native entry) with the emulated library behind the binding.  The batch-parity
    tests over whole fixtures are left to the GPU (minutes per shape here), and the
    dantzig.rust extension links the real library, so it is not part of this run."""
    lib = _build()
    env = dict(os.environ, DZ_LIB=lib, DZ_LIB_TEST_ONLY="1")
    r = subprocess.run(
        [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q",
         "-x", "-p", "no:cacheprovider", "-k",
         "not test_full_config2_properties and not test_batch_parity and not test_batch_order "
         "and not rust_module and not through_the_module and not solve_batch_multi"],
        env=env, capture_output=True, text=True, timeout=1200, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:]
    assert " passed" in r.stdout and "failed" not in r.stdout
