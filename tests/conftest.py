import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        from dantzig_b200 import device_count

        return device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a device should fail loudly, not skip: the
    # product has no CPU path.  Without -m, gpu tests are skipped on CPU boxes.
    if config.getoption("-m"):
        return
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import dzo_py

    dzo_py.build()
    return dzo_py


def reference_frontend_dir():
    """Directory that holds the reference's unmodified `dantzig` Python package, or None:
    $DANTZIG_FRONTEND, the reference checkout (this container), or the copy that
    __graft_entry__.build() installs under baseline/_ref/ (what the GPU box has)."""
    for cand in (os.environ.get("DANTZIG_FRONTEND"), "/root/reference/python-source",
                 os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "dantzig", "optimize.py")):
            return cand
    return None
