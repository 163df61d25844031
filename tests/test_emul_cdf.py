"""csrc/cdf_core.h (the gather form of scoreCDF's dynamic programme that cdf_kernel runs), compiled for the host,
against the oracle's literal scatter loop (Motif.hs:282-326): the compressed CDF must be bit-identical."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from bioinformatics_toolkit_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul_cdf():
    so = os.path.join(HERE, "emul", "libemul.so")
    src = os.path.join(HERE, "emul", "emul_scan.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-w", "-shared", "-o", so, src])
    L = C.CDLL(so)
    L.emul_score_cdf.restype = C.c_int64

    def run(bg, mat):
        mat = np.ascontiguousarray(mat, dtype=np.float64)
        bga = np.array(bg, dtype=np.float64)
        xs, ps = np.zeros(200001), np.zeros(200001)
        vp = C.c_void_p
        n = L.emul_score_cdf(bga.ctypes.data_as(vp), C.c_int(mat.shape[0]), mat.ctypes.data_as(vp), xs.ctypes.data_as(vp), ps.ctypes.data_as(vp),
                             C.c_int64(200001))
        return xs[:n], ps[:n]
    return run


def _same(a, b):
    return a.size == b.size and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def test_fixture_motifs_bit_identical(emul_cdf, oracle, ref_motifs):
    bg = (0.25, 0.25, 0.25, 0.25)
    for name, mat in ref_motifs[:4]:
        x, p = emul_cdf(bg, mat)
        ox, op = oracle.score_cdf(bg, mat)
        assert _same(x, ox) and _same(p, op), name


def test_synthetic_shapes_and_backgrounds(emul_cdf, oracle):
    mats = synth.synth_pwms(6, 31, nmin=1, nmax=22) + synth.config_c4_motifs(800)[::160]
    mats.append(np.array([[0.25] * 4] * 5))                                   # lo == hi columns are skipped (Motif.hs:299)
    mats.append(np.array([[1.0, 0, 0, 0], [0, 0, 0, 1.0], [0, 0.5, 0.5, 0], [0.97, 0.01, 0.01, 0.01]]))
    # equal entries in a column: the runs of two bases coincide and must be merged element by element in (x, base) order
    mats.append(np.array([[0.4, 0.4, 0.1, 0.1], [0.1, 0.2, 0.2, 0.5], [0.3, 0.3, 0.3, 0.1], [0.25, 0.25, 0.3, 0.2]] * 3))
    for bg in ((0.25, 0.25, 0.25, 0.25), (0.3, 0.2, 0.2, 0.3), (0.1, 0.4, 0.15, 0.35)):
        for i, mat in enumerate(mats):
            x, p = emul_cdf(bg, mat)
            ox, op = oracle.score_cdf(bg, mat)
            assert _same(x, ox) and _same(p, op), (bg, i)
