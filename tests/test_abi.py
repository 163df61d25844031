"""The C-ABI shared library loads and exports every symbol include/pwmscan.h declares; without a
CUDA device it fails loudly (no CPU fallback).  No compute calls here."""
import ctypes
import os

import pytest

from bioinformatics_toolkit_b200 import _lib
from tests.conftest import HAS_GPU


def test_library_is_built():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    assert os.path.exists(_lib.LIB_PATH)


def test_exports_every_declared_symbol():
    L = _lib.load()
    names = _lib.exported_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libpwmscan.so does not export %s" % n
    assert L.pwmscan_abi_version() == 1


@pytest.mark.skipif(HAS_GPU, reason="checks the no-device behaviour")
def test_fails_loudly_without_a_device():
    with pytest.raises(_lib.PwmScanError) as e:
        _lib.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.dirname(_lib.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
