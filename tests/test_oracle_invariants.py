"""The oracle against everything the reference's own tests hold for this path:
the five invariants of bioinformatics-toolkit/tests/Tests/Motif.hs:58-97 over tests/data/motifs.fasta,
plus the oracle's frozen goldens.  (The reference's test DNA — randomRs (0,3) (mkStdGen 2) — depends
on the `random` package and cannot be regenerated without GHC; a SplitMix64 stream stands in.)
"""
import json
import os

import numpy as np
import pytest

from bioinformatics_toolkit_b200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
BG = (0.25, 0.25, 0.25, 0.25)


@pytest.fixture(scope="module")
def dna():
    return synth.synth_dna(5000, 2)


def test_findTFBS_equals_findTFBSSlow(oracle, ref_motifs, dna):
    # findTFBSTest, Tests/Motif.hs:58-64: exact equality of [(Int, Double)]
    pwm = ref_motifs[0][1]
    th = 0.6 * oracle.optimal_score(BG, pwm)
    for mode in (0, 1):
        a = oracle.find_tfbs(BG, pwm, dna, th, True, mode=mode)
        b = oracle.find_tfbs_slow(BG, pwm, dna, th)
        assert len(a[0]) > 0
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_maxMatchSc_equals_max_scores(oracle, ref_motifs, dna):
    # maxScTest, Tests/Motif.hs:66-71
    for _, pwm in ref_motifs:
        assert oracle.max_match_sc(BG, pwm, dna) == oracle.scores(BG, pwm, dna).max()


def test_pValueToScore_vs_exact(oracle, ref_motifs):
    # pValueTest, Tests/Motif.hs:73-80: round (10 * x) equal
    for _, pwm in ref_motifs:
        assert round(10 * oracle.pvalue_to_score(1e-4, BG, pwm)) == round(10 * oracle.pvalue_to_score_exact(1e-4, BG, pwm))


def test_cdf_truncate(oracle, ref_motifs):
    # cdfTruncateTest, Tests/Motif.hs:82-89: exact
    for _, pwm in ref_motifs:
        c = oracle.score_cdf(BG, pwm)
        assert oracle.cdf_inv(c, 1 - 1e-4) == oracle.cdf_inv(oracle.truncate_cdf(0.999, c), 1 - 1e-4)


def test_reverse_complement(oracle, ref_motifs, dna):
    # reverseComplentary, Tests/Motif.hs:91-97: equal after round (1e5 * x)
    rcdna = np.frombuffer(bytes(dna)[::-1].translate(bytes.maketrans(b"ACGT", b"TGCA")), dtype=np.uint8)
    for _, pwm in ref_motifs:
        s1 = oracle.scores(BG, pwm, dna)
        s2 = oracle.scores(BG, oracle.rc_pwm(pwm), rcdna)[::-1]
        assert np.array_equal(np.round(1e5 * s1), np.round(1e5 * s2))


def test_frozen_goldens(oracle, ref_motifs, ref_motifs_meme, dna):
    g = json.load(open(os.path.join(GOLDEN, "oracle_goldens.json")))
    for (name, pwm), e in zip(ref_motifs, g["fasta"]):
        assert name == e["name"]
        assert oracle.pvalue_to_score(1e-4, BG, pwm).hex() == e["pvalue_to_score_1e-4"]
        xs, ps = oracle.score_cdf(BG, pwm)
        assert xs.size == e["cdf_points"] and xs[-1].hex() == e["cdf_x_last"] and ps[0].hex() == e["cdf_p_first"]
        assert [v.hex() for v in oracle.optimal_scores_suffix(BG, pwm)] == e["sigma"]
        pos, sc = oracle.find_tfbs(BG, pwm, dna, 0.6 * oracle.optimal_score(BG, pwm), True)
        assert pos.tolist() == e["tfbs_pos"] and [v.hex() for v in sc] == e["tfbs_sc"]
    mats = [m for _, m in ref_motifs] + [m for _, m in ref_motifs_meme]
    for i, j, d, sd, p in g["align"]:
        dd, (ss, pp) = oracle.alignment(mats[i], mats[j])
        assert (dd.hex(), ss, pp) == (d, sd, p)


def test_survey_smoke_values(oracle, ref_motifs):
    # values from an independent throw-away restatement (Python) of Q5/Q6
    exp = [7.009685376479372, 7.226297856166506, 7.118708444675251, 7.139596009320544, 7.35050148897909]
    npts = [151394, 153842, 134835, 152244, 112309]
    for (_, pwm), v, n in zip(ref_motifs, exp, npts):
        assert oracle.pvalue_to_score(1e-4, BG, pwm) == v
        assert oracle.score_cdf(BG, pwm)[0].size == n
    m1, m2 = ref_motifs[0][1], ref_motifs[1][1]
    assert oracle.alignment(m1, m1) == (0.0, (True, 0))
    assert oracle.alignment(m1, oracle.rc_pwm(m1)) == (0.0, (False, 0))
    assert oracle.alignment(m1, m2) == (0.12889355547153183, (True, 1))


def test_skip_and_iupac_semantics(oracle, ref_motifs):
    pwm = ref_motifs[2][1]
    dna = bytearray(bytes(synth.synth_dna(2000, 5)))
    dna[100:110] = b"N" * 10
    dna[500] = ord("R")
    th = 0.5 * oracle.optimal_score(BG, pwm)
    pos, _ = oracle.find_tfbs(BG, pwm, bytes(dna), th, True)
    n = pwm.shape[0]
    assert not any((p <= 109 and p + n > 100) or (p <= 500 < p + n) for p in pos)   # windows touching non-ACGT never hit
    # IUPAC-aware search equals the brute force (N scores 0)
    a = oracle.find_tfbs(BG, pwm, bytes(dna), th, False)
    b = oracle.find_tfbs_slow(BG, pwm, bytes(dna), th)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    dna[700] = ord("X")
    with pytest.raises(ValueError):
        oracle.scores(BG, pwm, bytes(dna))


def test_short_and_empty(oracle, ref_motifs):
    pwm = ref_motifs[0][1]
    assert oracle.find_tfbs(BG, pwm, b"ACGT", 0.0, True)[0].size == 0
    assert oracle.find_tfbs(BG, pwm, b"", 0.0, True)[0].size == 0
    assert oracle.scores(BG, pwm, b"ACG").size == 0
    assert oracle.max_match_sc(BG, pwm, b"ACG") == float("-inf")


def test_to_affinity(oracle):
    # Data/Bed/Utils.hs:120-123
    assert oracle.to_affinity(1e-5) == 500 and oracle.to_affinity(0.5) == 9 and oracle.to_affinity(0.0) == 1000


def test_app_oracle_mt_composition_equals_plain(oracle):
    """oracle/app_oracle.scan_motif_lines_mt (one multi-threaded oracle scan per region) prints exactly the lines of the
    plain per-(region, motif, strand) restatement of scanMotif (Data/Bed/Utils.hs:101-124)."""
    from oracle import app_oracle
    gen = {b"chr1": synth.synth_dna(30_000, 31, gc=0.45, lower_frac=0.2, n_runs=[(9000, 300)]), b"chr2": synth.synth_dna(8_000, 32)}
    mats = synth.synth_pwms(9, 44, nmin=6, nmax=16)
    names = [b"M%d" % i for i in range(len(mats))]
    rng = np.random.default_rng(2)
    beds = [(c, int(s), int(s) + int(w)) for c, s, w in zip(rng.choice([b"chr1", b"chr2", b"chrQ"], 40), rng.integers(0, 7800, 40), rng.integers(1, 600, 40))]
    a = app_oracle.scan_motif_lines(gen, names, mats, beds, p=1e-3)
    b = app_oracle.scan_motif_lines_mt(gen, names, mats, beds, p=1e-3, nthreads=4)
    assert a == b and len(a) > 30
