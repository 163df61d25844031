"""Bio.Motif.Search mirror (bioinformatics-toolkit/src/Bio/Motif/Search.hs).

findTFBS (25-36), findTFBSWith (38-58), findTFBSSlow (87-95), maxMatchSc (98-109),
optimalScoresSuffix (111-124).  The conduits of (Int, Double) become Python lists / arrays of
(position, score) in ascending position; everything is computed by the CUDA library.
"""
import numpy as np

from . import _lib
from .motif import _dev_motifs, _dev_seq, scores, size


def optimalScoresSuffix(bg, pwm):
    return _dev_motifs(bg, [pwm]).get_sigma(0, 0)


def findTFBSWith(sigma, bg, pwm, dna, thres, skip=True):
    """[(i, score)] of the windows passing the look-ahead search with the caller's sigma."""
    if isinstance(dna, _lib.DeviceSeq):
        mo = _dev_motifs(bg, [pwm], dna.ctx)
    else:
        mo = _dev_motifs(bg, [pwm])
    if sigma is not None:
        mo.set_sigma(0, 0, sigma)
    plan = _lib.Plan(mo, [thres], strands=1, skip_ambiguous=skip)
    if isinstance(dna, _lib.DeviceSeq):
        hits = plan.scan(dna)
    else:
        hits = plan.scan_host(dna.bs if hasattr(dna, "bs") else bytes(dna))
    return list(zip(hits.pos.tolist(), hits.score.tolist()))


def findTFBS(bg, pwm, dna, thres, skip=True):
    return findTFBSWith(None, bg, pwm, dna, thres, skip)


def findTFBSSlow(bg, pwm, dna, thres):
    """scores' filtered by (>= thres) (Search.hs:87-95)."""
    sc = scores(bg, pwm, dna)
    idx = np.nonzero(sc >= thres)[0]
    return list(zip(idx.tolist(), sc[idx].tolist()))


def maxMatchSc(bg, pwm, dna):
    seq = _dev_seq(dna, keep_ascii=True)
    if seq.length < size(pwm):
        return float("-inf")
    return float(_lib.max_match_sc(seq, _dev_motifs(bg, [pwm], seq.ctx))[0])
