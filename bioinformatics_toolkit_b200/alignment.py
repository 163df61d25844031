"""Bio.Motif.Alignment mirror (bioinformatics-toolkit/src/Bio/Motif/Alignment.hs).

An AlignFn in the reference is an arbitrary Haskell closure; the combinations the package ships
(jsd x {lin,quad,cub,exp}Penal x {l1,l2,l3,lInf}; app/MergeMotifs.hs:129-144) are lowered to the
CUDA library.  An `AlignFn` object here is callable on two PWMs (-> (d, (isSame, pos))) and also
exposes the batched device entry points `all_pairs` / `pairs` the merge code uses.
"""
import numpy as np

from . import _lib


class AlignFn:
    def __init__(self, penal="exp", gap=0.05, comb="l1"):
        if penal not in _lib.PENAL:
            raise ValueError("Unknown gap mode")
        if comb not in _lib.COMB:
            raise ValueError("Unknown average mode")
        self.penal, self.gap, self.comb = penal, float(gap), comb

    def _motifs(self, pwms, ctx=None):
        return _lib.DeviceMotifs(ctx or _lib.default_context(), [p._mat for p in pwms])

    def __call__(self, m1, m2):
        d, sd, pos = self._motifs([m1, m2]).align_pairs([0], [1], self.penal, self.gap, self.comb)
        return float(d[0]), (bool(sd[0]), int(pos[0]))

    def all_pairs(self, pwms):
        """align pwm_i pwm_j for all i < j, packed upper triangle in row-major order."""
        return self._motifs(pwms).align_all_pairs(self.penal, self.gap, self.comb)

    def pairs(self, pwms, a, b):
        return self._motifs(pwms).align_pairs(a, b, self.penal, self.gap, self.comb)

    def resident(self, pwms, ctx=None):
        """The PWMs kept on the device for a merge loop (`_lib.AlignSet`): `.pairs()` = all i < j, `.update(slot, mat)`,
        `.pairs(a, b)` = row refresh."""
        return _lib.AlignSet(ctx or _lib.default_context(), [p._mat for p in pwms], self.penal, self.gap, self.comb)


def alignmentBy(distance="jsd", penal=("exp", 0.05), comb="l1"):
    """alignmentBy jsd (pFn gap) combFn (Alignment.hs:83-86).  Only the Jensen-Shannon distance the
    package ships is available on the device."""
    if distance != "jsd":
        raise ValueError("only the jsd DistanceFn is lowered to the GPU")
    return AlignFn(penal[0], penal[1], comb)


def linPenal(x):
    return ("linear", x)


def quadPenal(x):
    return ("quadratic", x)


def cubPenal(x):
    return ("cubic", x)


def expPenal(x):
    return ("exp", x)


l1, l2, l3, lInf = "l1", "l2", "l3", "max"

alignment = AlignFn("exp", 0.05, "l1")      # alignment = alignmentBy jsd (expPenal 0.05) l1 (Alignment.hs:43-44)


def pair_index(i, j, M):
    """Position of pair (i < j) in the packed upper triangle."""
    return i * M - i * (i + 1) // 2 + (j - i - 1)
