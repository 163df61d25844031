"""Text -> Double as the reference does it (Bio.Utils.Misc.readDouble = bytestring-lexing's `readSigned readExponential`,
Utils/Misc.hs:21-25), pValueToScoreExact (Motif.hs:218-230) and the `%f` of toMEME's header (Motif.hs:351): host-side
restatements in bioinformatics_toolkit_b200/motif.py.  PARITY STATUS: the third-party algorithm is restated from its
published source (not under /root/reference) — unpinned against a GHC build."""
import json
import os
import random

import numpy as np
import pytest

from bioinformatics_toolkit_b200 import motif as Mo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_reference_fixture_tokens():
    """The number tokens of the reference's own motif fixtures (tests/data/motifs.{fasta,meme}).  Most have <= 16 significant
    digits: one correctly rounded division of two exact operands, the same double as strtod.  Fifteen carry 17 digits
    (e.g. 0.010290750000000001): their mantissa exceeds 2^53, fromInteger rounds it, the division rounds again, and the
    reference's PWM entry is ONE ULP away from strtod's — the committed goldens (ref_motifs_*.json) hold the reference's."""
    toks = json.load(open(os.path.join(GOLDEN, "ref_motif_tokens.json")))["tokens"]
    assert len(toks) == 221
    diff = [t for t in toks if Mo.read_double(t.encode()) != float(t)]
    assert diff == ['0.010290750000000001', '0.012137499999999999', '0.012410250000000001', '0.013327499999999999',
                    '0.014602750000000001', '0.022783499999999998', '0.028267999999999998', '0.038650500000000004',
                    '0.052595249999999996', '0.052962999999999996', '0.053024999999999996', '0.13376949999999999',
                    '0.20135799999999998', '0.22489399999999998', '0.38844975000000004']
    for t in toks:
        a, b = Mo.read_double(t.encode()), float(t)
        assert abs(a - b) <= np.spacing(b) and (a == b or len(t.replace("0.", "").lstrip("0")) >= 17)
    # the goldens were parsed with read_double
    fa = json.load(open(os.path.join(GOLDEN, "ref_motifs_fasta.json")))["motifs"]
    vals = {v for m in fa for row in m["rows"] for v in row}
    assert Mo.read_double(b"0.010290750000000001") in vals and float("0.010290750000000001") not in vals


def test_grammar():
    rd = Mo.read_double
    assert rd(b"1") == 1.0 and rd(b"1.") == 1.0 and rd(b"12.5abc") == 12.5 and rd(b"1e") == 1.0 and rd(b"1e+") == 1.0
    assert rd(b"+2.5E2") == 250.0 and rd(b"-0.5") == -0.5 and rd(b"1e-3") == 0.001 and rd(b"007.250") == 7.25
    assert rd(b"-0.0") == 0.0 and np.signbit(rd(b"-0.0")) and rd(b"0.000") == 0.0 and rd(b"0e999999") == 0.0
    assert rd(b"3.2e-035") == 3.2e-35
    for bad in (b"", b".5", b"-", b"+", b"abc", b"-.5", b"nan", b"inf", b" 1"):
        with pytest.raises(ValueError):
            rd(bad)


def test_not_correctly_rounded_where_the_library_is_not():
    """Mantissas above 2^53 are rounded twice (fromInteger, then the division) and powers of ten above 10^22 are computed
    by repeated squaring in Double: known cases where readDouble and strtod disagree."""
    rd = Mo.read_double
    known = {b"0.497869073662585178": None, b"0.7298063513969371648": None, b"1e-23": None}
    for t in known:
        a, b = rd(t), float(t)
        assert a != b and abs(a - b) <= 2 * np.spacing(b), t
    assert rd(b"5e-324") == 0.0                      # 5 / 10^324 = 5 / inf
    assert rd(b"1e23") == 1e23 and rd(b"123456789012345678901234567890") == 1.2345678901234568e29
    # below 2^53 and scale <= 22: always strtod's answer
    rng = random.Random(7)
    for _ in range(20000):
        nd = rng.randint(1, 15)
        s = ("0." + "".join(rng.choice("0123456789") for _ in range(nd))).encode()
        assert rd(s) == float(s)
    # above: a sizeable fraction differs (double rounding), never by more than one ulp
    ndiff = 0
    for _ in range(20000):
        s = ("0." + "".join(rng.choice("0123456789") for _ in range(rng.randint(17, 19)))).encode()
        a, b = rd(s), float(s)
        ndiff += a != b
        assert abs(a - b) <= np.spacing(b)
    assert 1000 < ndiff < 15000


def test_pow10_squaring():
    for k in range(0, 23):
        assert Mo._pow10(k) == float(10 ** k)
    assert Mo._pow10(23) == 1e23 or abs(Mo._pow10(23) - 1e23) <= np.spacing(1e23)
    assert Mo._pow10(324) == float("inf")


def test_to_pwm_and_meme_go_through_read_double():
    pwm = Mo.toPWM([b"0.25 0.25 0.25 0.25", b"1e-1 0.7298063513969371648 0.1 0.0701936486030628352"])
    assert pwm._mat[1, 1] == Mo.read_double(b"0.7298063513969371648") != float("0.7298063513969371648")
    meme = b"MEME version 4\n\nMOTIF m1\nletter-probability matrix: alength= 4 w= 1 nsites= 3 E= 0\n 0.497869073662585178 0.1 0.2 0.2\n\n"
    m = Mo.fromMEME(meme)[0]
    assert m._pwm._mat[0, 0] == Mo.read_double(b"0.497869073662585178") and m._pwm._nSites == 3


def test_meme_header_uses_show_ffloat():
    """printf "%f" without a precision is showFFloat Nothing: 'A 0.25 C 0.25 ...', not C's 'A 0.250000 ...'."""
    sf = Mo._show_ffloat
    assert [sf(v) for v in (0.25, 1.0, 0.0, 1e-7, 0.1 + 0.2, 123.5, 1e22, -0.5, 0.303)] == \
        ["0.25", "1.0", "0.0", "0.0000001", "0.30000000000000004", "123.5", "10000000000000000000000.0", "-0.5", "0.303"]
    txt = Mo.toMEME([Mo.Motif(b"x", Mo.PWM(None, [[0.25, 0.25, 0.25, 0.25]]))])
    assert txt.startswith(b"MEME version 4\n\nALPHABET= ACGT\n\nstrands: + -\n\nBackground letter frequencies\nA 0.25 C 0.25 G 0.25 T 0.25\n\nMOTIF x\n")
    txt = Mo.toMEME([], Mo.Bkgd(0.303, 0.183, 0.209, 0.306))
    assert b"A 0.303 C 0.183 G 0.209 T 0.306\n" in txt


def test_pvalue_to_score_exact_matches_oracle():
    from oracle import oracle as O
    from bioinformatics_toolkit_b200 import synth
    ref = json.load(open(os.path.join(GOLDEN, "ref_motifs_fasta.json")))["motifs"]
    mats = [np.array(m["rows"]) for m in ref] + synth.synth_pwms(4, 3, nmin=5, nmax=9)
    for bg in (Mo.default_bkgd, Mo.Bkgd(0.3, 0.2, 0.2, 0.3)):
        for mat in mats:
            for p in (1e-4, 1e-3, 0.05):
                a = Mo.pValueToScoreExact(p, bg, Mo.PWM(None, mat))
                b = O.pvalue_to_score_exact(p, bg.freqs, mat)
                assert a.hex() == float(b).hex()
    # the reference's own pValueTest (tests/Tests/Motif.hs:73-84) pins it within 1e-2 relative of pValueToScore: SURVEY 8(c) smoke values
    smoke = [7.0099, 7.2266, 7.1188, 7.1395, 7.3509]
    for mat, v in zip(mats[:5], smoke):
        assert abs(Mo.pValueToScoreExact(1e-4, Mo.default_bkgd, Mo.PWM(None, mat)) - v) < 1e-3
    with pytest.raises(IndexError):
        Mo.pValueToScoreExact(1.0, Mo.default_bkgd, Mo.PWM(None, mats[-1]))
