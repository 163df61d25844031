"""csrc/align_core.h (the function the alignment kernel runs, here compiled for the host with g++)
against the oracle's independent restatement of alignmentBy: bit-identical distance, direction and
position for every shipped penalty / combiner."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from bioinformatics_toolkit_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul_align():
    so = os.path.join(HERE, "emul", "libemul.so")
    src = os.path.join(HERE, "emul", "emul_scan.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-w", "-shared", "-o", so, src])
    L = C.CDLL(so)

    def run(m1, m2, penal, gap, comb):
        from oracle.oracle import PENAL, COMB
        m1 = np.ascontiguousarray(m1, dtype=np.float64)
        m2 = np.ascontiguousarray(m2, dtype=np.float64)
        d, sd, pos, amb = C.c_double(0), C.c_int(0), C.c_int(0), C.c_int(0)
        vp = C.c_void_p
        L.emul_align(C.c_int(m1.shape[0]), m1.ctypes.data_as(vp), C.c_int(m2.shape[0]), m2.ctypes.data_as(vp), C.c_int(PENAL[penal]),
                     C.c_double(gap), C.c_int(COMB[comb]), C.byref(d), C.byref(sd), C.byref(pos), C.byref(amb))
        return d.value, (bool(sd.value), pos.value), amb.value
    return run


def align_test_motifs(oracle):
    mats = synth.synth_pwms(24, 15, nmin=4, nmax=22)
    mats.append(oracle.rc_pwm(mats[3]))                       # reverse-complement duplicate
    mats.append(mats[5].copy())                               # exact duplicate
    pal = mats[7][:6]
    mats.append(np.vstack([pal, oracle.rc_pwm(pal)]))         # exact palindrome
    mats.append(np.array([[0.25] * 4] * 7))
    mats.append(np.array([[1.0, 0, 0, 0], [0, 0, 0, 1.0], [0, 0.5, 0.5, 0]]))   # exact zeros
    return mats


@pytest.mark.parametrize("penal,comb,gap", [("exp", "l1", 0.05), ("linear", "l2", 0.1), ("quadratic", "l3", 0.02),
                                           ("cubic", "max", 0.05), ("exp", "max", 0.3)])
def test_host_build_equals_oracle(emul_align, oracle, penal, comb, gap):
    mats = align_test_motifs(oracle)
    namb = 0
    for i, a in enumerate(mats):
        for j, b in enumerate(mats):
            d, dp, amb = emul_align(a, b, penal, gap, comb)
            od, odp = oracle.alignment(a, b, penal, gap, comb)
            assert (d.hex(), dp) == (od.hex(), odp), (i, j)
            namb += amb
    assert namb < len(mats) ** 2 // 10       # structural ties (duplicates, palindromes) are bit-equal, hence not flagged
