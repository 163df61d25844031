import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import ctypes
        lib = ctypes.CDLL("libcuda.so.1")
        if lib.cuInit(0) != 0:
            return False
        n = ctypes.c_int(0)
        return lib.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def ref_motifs():
    """PWMs of the reference's own fixture tests/data/motifs.fasta (see tests/golden/make_golden.py)."""
    d = json.load(open(os.path.join(GOLDEN, "ref_motifs_fasta.json")))
    return [(m["name"], np.array(m["rows"], dtype=np.float64)) for m in d["motifs"]]


@pytest.fixture(scope="session")
def ref_motifs_meme():
    d = json.load(open(os.path.join(GOLDEN, "ref_motifs_meme.json")))
    return [(m["name"], np.array(m["rows"], dtype=np.float64)) for m in d["motifs"]]


@pytest.fixture(scope="session")
def ctx():
    from bioinformatics_toolkit_b200 import _lib
    return _lib.default_context(0)
