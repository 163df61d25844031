"""Bio.Seq mirror (bioinformatics-toolkit/src/Bio/Seq.hs): DNA with a phantom alphabet.

`DNA` wraps upper-cased bytes (Seq.hs:41); `fromBS` upper-cases and validates against the alphabet
(Seq.hs:86-95); `rc` is the reverse complement (Seq.hs:111-119).  Host-side glue: the heavy lifting
(packing, validation of genome-scale inputs) happens on the device in `_lib.DeviceSeq`.
"""
import numpy as np

ALPHABET = {"Basic": b"ACGT", "IUPAC": b"ACGTNVHDBMKWSYR"}
_RC = bytes.maketrans(b"ACGT", b"TGCA")
_UPPER = np.arange(256, dtype=np.uint8)
_UPPER[ord("a"):ord("z") + 1] -= 32


class DNA:
    __slots__ = ("bs", "alphabet")

    def __init__(self, bs, alphabet="IUPAC"):
        self.bs = bytes(bs)
        self.alphabet = alphabet

    def __len__(self):
        return len(self.bs)

    def __repr__(self):
        return self.bs.decode("latin1")


def toBS(dna):
    return dna.bs


def unsafeFromBS(bs, alphabet="IUPAC"):
    return DNA(bs, alphabet)


def length(dna):
    return len(dna.bs)


def slice(i, l, dna):  # noqa: A001  (Seq.hs:72)
    return DNA(dna.bs[i:i + l], dna.alphabet)


def fromBS(bs, alphabet="IUPAC"):
    """Right DNA | Left message, as (ok, value) — Seq.hs:86-95."""
    arr = _UPPER[np.frombuffer(bytes(bs), dtype=np.uint8)]
    ok = np.zeros(256, dtype=bool)
    ok[np.frombuffer(ALPHABET[alphabet], dtype=np.uint8)] = True
    bad = np.nonzero(~ok[arr])[0]
    if bad.size:
        return False, "Bio.Seq.fromBS: unknown character: " + chr(int(arr[bad[0]]))
    return True, DNA(arr.tobytes(), alphabet)


def rc(dna):
    return DNA(dna.bs[::-1].translate(_RC), dna.alphabet)
