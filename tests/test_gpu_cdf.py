"""Device scoreCDF / cdf' / pValueToScore (Motif.hs:213-326) vs the oracle: bit-identical."""
import json
import os

import numpy as np
import pytest

from bioinformatics_toolkit_b200 import _lib, synth

pytestmark = pytest.mark.gpu
BG = (0.25, 0.25, 0.25, 0.25)
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_score_cdf_points_bit_identical(ctx, oracle, ref_motifs, ref_motifs_meme):
    mats = [m for _, m in ref_motifs] + [m for _, m in ref_motifs_meme]
    mo = _lib.DeviceMotifs(ctx, mats, BG)
    for m, mat in enumerate(mats):
        xs, ps = mo.score_cdf(m)
        ox, op = oracle.score_cdf(BG, mat)
        assert xs.size == ox.size
        assert np.array_equal(xs.view(np.uint64), ox.view(np.uint64))
        assert np.array_equal(ps.view(np.uint64), op.view(np.uint64))


def test_pvalue_to_score_fixture_goldens(ctx, oracle, ref_motifs):
    g = json.load(open(os.path.join(GOLDEN, "oracle_goldens.json")))
    mo = _lib.DeviceMotifs(ctx, [m for _, m in ref_motifs], BG)
    th = mo.pvalue_to_score(1e-4)
    assert [float(v).hex() for v in th] == [e["pvalue_to_score_1e-4"] for e in g["fasta"]]
    # pValueTest (Tests/Motif.hs:73-80) through the device path
    for (_, mat), v in zip(ref_motifs, th):
        assert round(10 * v) == round(10 * oracle.pvalue_to_score_exact(1e-4, BG, mat))


def test_config_thresholds_known_answers(ctx):
    d = json.load(open(os.path.join(GOLDEN, "config_thresholds.json")))
    for key, mats in (("c1", [synth.ctcf_like_pwm()]), ("c2", synth.synth_pwms(100, 12)), ("c4", synth.config_c4_motifs(800))):
        th = _lib.DeviceMotifs(ctx, mats, BG).pvalue_to_score(1e-5)
        assert [float(v).hex() for v in th] == d[key]


def test_nonuniform_background_long_motif_and_errors(ctx, oracle):
    bg = (0.3, 0.2, 0.2, 0.3)
    mats = synth.synth_pwms(6, 91, nmin=3, nmax=40) + [np.array([[0.25] * 4] * 5), np.array([[1.0, 0, 0, 0]] * 4)]
    mo = _lib.DeviceMotifs(ctx, mats, bg)
    for p in (1e-3, 1e-6, 0.5):
        th = mo.pvalue_to_score(p)
        for m, mat in enumerate(mats):
            try:
                ref = oracle.pvalue_to_score(p, bg, mat)
            except ValueError:
                continue
            assert float(th[m]).hex() == ref.hex()
    with pytest.raises(_lib.PwmScanError):
        mo.pvalue_to_score(-0.5)
