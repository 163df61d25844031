"""Device alignmentBy (Motif/Alignment.hs:83-132) vs the oracle: (same_dir, pos) identical,
distance within 1e-12 relative (the device evaluates log of the mixture columns with the CUDA
libm; pairs with near-tied decisions are re-evaluated on the host and are bit-identical)."""
import numpy as np
import pytest

from bioinformatics_toolkit_b200 import _lib, synth
from tests.test_emul_align import align_test_motifs

pytestmark = pytest.mark.gpu
BG = (0.25, 0.25, 0.25, 0.25)
TOL = 1e-12


def close(a, b):
    return np.all(np.abs(a - b) <= TOL * np.maximum(1e-300, np.maximum(np.abs(a), np.abs(b))))


@pytest.mark.parametrize("penal,comb,gap", [("exp", "l1", 0.05), ("linear", "l2", 0.1), ("quadratic", "l3", 0.02), ("cubic", "max", 0.05)])
def test_all_pairs_and_pair_list(ctx, oracle, penal, comb, gap):
    mats = align_test_motifs(oracle)
    mo = _lib.DeviceMotifs(ctx, mats, BG)
    d, sd, pos = mo.align_all_pairs(penal, gap, comb)
    od, osd, opos = oracle.alignment_all_pairs(mats, penal, gap, comb)
    assert np.array_equal(sd, osd) and np.array_equal(pos, opos)
    assert close(d, od)
    # explicit list, both argument orders (Merge.hs:142-143 relies on the order)
    M = len(mats)
    a = np.repeat(np.arange(M), M).astype(np.int32)
    b = np.tile(np.arange(M), M).astype(np.int32)
    d2, sd2, pos2 = mo.align_pairs(a, b, penal, gap, comb)
    for k in range(0, M * M, 7):
        rd, (rs, rp) = oracle.alignment(mats[a[k]], mats[b[k]], penal, gap, comb)
        assert (bool(sd2[k]), int(pos2[k])) == (rs, rp)
        assert abs(d2[k] - rd) <= TOL * max(abs(rd), 1e-300)


def test_config_c5_sample(ctx, oracle):
    """BASELINE config 5: 2 000 motifs all-vs-all (1 999 000 pairs); a sample of rows against the oracle."""
    mats = synth.config_c5_motifs(2000)
    mo = _lib.DeviceMotifs(ctx, mats, BG)
    d, sd, pos = mo.align_all_pairs()
    M = len(mats)
    assert d.size == M * (M - 1) // 2 and np.all(np.isfinite(d)) and np.all(d >= -1e-12)
    rng = np.random.default_rng(0)
    for i in rng.choice(M - 1, size=12, replace=False):
        base = i * M - i * (i + 1) // 2
        sub = [mats[i]] + mats[i + 1:]
        # oracle for row i only: pairs (i, j), j > i
        for j in range(i + 1, min(M, i + 400)):
            rd, (rs, rp) = oracle.alignment(mats[i], mats[j])
            k = base + (j - i - 1)
            assert (bool(sd[k]), int(pos[k])) == (rs, rp), (i, j)
            assert abs(d[k] - rd) <= TOL * max(abs(rd), 1e-300)
