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


def test_bed6_formatter_host_only():
    """pwmscan_format_bed6 is pure host code (toLine of mkBed): exact text for random hits, negative numbers,
    multi-digit boundaries; threads write into one exactly-sized buffer."""
    import ctypes as C
    import numpy as np
    L = C.CDLL(_lib.LIB_PATH)
    n = 150000
    rng = np.random.default_rng(1)
    seg = rng.integers(0, 50, n).astype(np.int32)
    mot = rng.integers(0, 7, n).astype(np.int32)
    strand = rng.integers(0, 2, n).astype(np.uint8)
    pos = rng.integers(0, 10 ** 9, n).astype(np.int64)
    pos[:20] = [0, 9, 10, 99, 100, 999, 1000, 9999, 10 ** 4, 10 ** 5 - 1, 10 ** 5, 10 ** 6 - 1, 10 ** 6, 10 ** 7, 10 ** 8 - 1, 10 ** 8, 10 ** 9 - 1, 1, 2, 3]
    score = rng.integers(-5, 1001, n).astype(np.int64)
    chroms = [b"chr1", b"chrX_long_name", b"c"]
    rchrom = rng.integers(0, 3, 50).astype(np.int32)
    rstart = rng.integers(0, 10 ** 8, 50).astype(np.int64)
    rstart[0] = 0
    coff = np.concatenate([[0], np.cumsum([len(c) for c in chroms])]).astype(np.int64)
    names = [b"m%d name" % i for i in range(7)]
    moff = np.concatenate([[0], np.cumsum([len(c) for c in names])]).astype(np.int64)
    mlen = rng.integers(5, 30, 7).astype(np.int32)
    text, tlen = C.c_char_p(), C.c_int64(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)      # noqa: E731
    rc = L.pwmscan_format_bed6(C.c_int64(n), p(seg), p(mot), p(strand), p(pos), p(score), p(rchrom), p(rstart),
                               C.c_char_p(b"".join(chroms)), p(coff), C.c_char_p(b"".join(names)), p(moff), p(mlen),
                               C.byref(text), C.byref(tlen))
    assert rc == 0
    got = C.string_at(text, tlen.value).split(b"\n")
    L.pwmscan_free.argtypes = [C.c_void_p]
    L.pwmscan_free(text)
    assert got[-1] == b"" and len(got) == n + 1
    for k in list(range(40)) + list(rng.integers(0, n, 2000)):
        st = int(rstart[seg[k]] + pos[k])
        assert got[k] == b"%s\t%d\t%d\t%s\t%d\t%s" % (chroms[rchrom[seg[k]]], st, st + mlen[mot[k]], names[mot[k]], score[k], b"+" if strand[k] == 0 else b"-")


def test_dist_formatter_host_only():
    """pwmscan_format_dist (host code): `merge_motifs dist` lines with toShortest numbers, all-pairs and explicit
    pairs, equal to the Python restatement of the ECMAScript formatting rules — including tiny and huge values."""
    import time
    import numpy as np
    from bioinformatics_toolkit_b200.motif import _shortest
    rng = np.random.default_rng(5)
    names = [b"m%d|x" % i for i in range(400)]
    n = len(names) * (len(names) - 1) // 2
    d = rng.random(n) * 10.0 ** rng.integers(-9, 3, n)
    d[:12] = [0.0, 1.0, 1e-5, 1e-6, 1e-7, 1.5e-7, 0.1, 123.5, 1e21, 1e20, 0.12889355547153183, 5e-324]
    got = _lib.format_dist(names, d).split(b"\n")
    assert got[-1] == b"" and len(got) == n + 1
    k = 0
    for i in range(len(names)):
        for j in range(i + 1, len(names)):
            if k < 3000 or k % 97 == 0:
                assert got[k] == b"%s\t%s\t%s" % (names[i], names[j], _shortest(d[k]).encode()), (k, got[k])
            k += 1
    a = rng.integers(0, len(names), 5000).astype(np.int32)
    b = rng.integers(0, len(names), 5000).astype(np.int32)
    got = _lib.format_dist(names, d[:5000], a, b).split(b"\n")[:-1]
    assert got == [b"%s\t%s\t%s" % (names[x], names[y], _shortest(v).encode()) for x, y, v in zip(a, b, d[:5000])]
    big = rng.random(2000 * 1999 // 2)
    t0 = time.perf_counter()
    text = _lib.format_dist([b"motif_%d" % i for i in range(2000)], big)
    assert text.count(b"\n") == big.size and time.perf_counter() - t0 < 20


def test_haskell_shim_binds_every_entry_point():
    """haskell/Bio/Motif/CUDA.hs (the uncompiled drop-in binding) has a foreign import for every function the header
    declares, so the shim and the C ABI cannot drift apart unnoticed."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hs = open(os.path.join(root, "haskell", "Bio", "Motif", "CUDA.hs")).read()
    missing = [n for n in _lib.exported_symbols() if ('"%s"' % n) not in hs]
    assert missing == []


def test_context_outlives_its_handles_whatever_the_finaliser_order():
    """The cyclic garbage collector runs finalisers in arbitrary order (module globals captured by lambdas, interpreter
    shutdown): a Context whose __del__ comes first must not destroy the C context under the handles that still hand
    buffers back to its pools — it is destroyed by whoever goes last (seen as an abort at exit under MALLOC_PERTURB_)."""
    calls = []

    class FakeL:
        def pwmscan_ctx_destroy(self, h):
            calls.append("ctx")

    ctx = _lib.Context.__new__(_lib.Context)
    ctx.L, ctx.h, ctx._live, ctx._closing = FakeL(), _lib._vp(1), 0, False
    ctx._retain()
    ctx._retain()
    ctx.close()                      # e.g. Context.__del__ called first by the collector
    assert calls == []
    ctx._release()
    assert calls == []
    ctx._release()                   # the last handle
    assert calls == ["ctx"] and not ctx.h
    ctx.close()
    assert calls == ["ctx"]          # idempotent
    idle = _lib.Context.__new__(_lib.Context)
    idle.L, idle.h, idle._live, idle._closing = FakeL(), _lib._vp(1), 0, False
    idle.close()
    assert calls == ["ctx", "ctx"]   # nothing alive: destroyed at once
