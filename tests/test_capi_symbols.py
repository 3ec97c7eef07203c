"""The C-ABI library loads and exports every symbol include/dantzig_b200.h
declares (no compute calls: this runs without a GPU)."""
import os
import re

from dantzig_b200 import _capi

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                      "include", "dantzig_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dz_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(_capi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = _capi.lib()
    for name in declared_functions():
        assert getattr(lib, name) is not None, name
    assert lib.dz_version() == 100


def test_no_cpu_fallback_without_device():
    """On a box without a CUDA device a solve must fail loudly."""
    import pytest

    from dantzig_b200 import device_count, solve_model
    from dantzig_b200._capi import DzError
    from dantzig_b200.model import ModelBuilder

    if device_count() > 0:
        pytest.skip("a GPU is present")
    mb = ModelBuilder()
    x = mb.nonneg()
    mb.maximize([(-1.0, x)])
    with pytest.raises(DzError) as e:
        solve_model(mb.build())
    assert e.value.code == _capi.ERR_CUDA


def test_product_never_imports_oracle():
    """Nothing under dantzig_b200/ may reference oracle/ (grep-level check)."""
    pkg = os.path.join(os.path.dirname(HEADER), "..", "dantzig_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "dzo_" not in text and "from oracle" not in text and "import oracle" not in text, f
