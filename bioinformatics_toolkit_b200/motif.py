"""Bio.Motif mirror (bioinformatics-toolkit/src/Bio/Motif.hs), same names and argument meaning.

Types: PWM (Motif.hs:59-62), Motif (94-97), Bkgd (101-104).  Compute entry points go to the CUDA
library through `_lib` (no CPU fallback): scores / scores' / score (127-149), scoreCDF (282-326),
pValueToScore (213-214).  Host-side glue kept here exactly where the Haskell drop-in keeps it in
Haskell: rcPWM (77-78), optimalScore (152-162), cdf / cdf' / truncateCDF on an already-computed CDF
(236-274), MEME / motif-fasta parsing and printing (329-386).
"""
import math

import numpy as np

from . import _lib
from .seq import DNA


class Bkgd:
    """newtype Bkgd = BG (a, c, g, t); `def` = uniform 0.25 (Motif.hs:101-104)."""

    def __init__(self, a=0.25, c=0.25, g=0.25, t=0.25):
        self.freqs = (float(a), float(c), float(g), float(t))

    def __iter__(self):
        return iter(self.freqs)


BG = Bkgd
default_bkgd = Bkgd()


class PWM:
    """k x 4 position weight matrix (Motif.hs:59-62)."""

    def __init__(self, nSites, mat):
        self._nSites = nSites
        self._mat = np.ascontiguousarray(mat, dtype=np.float64).reshape(-1, 4)


class Motif:
    def __init__(self, name, pwm):
        self._name = name
        self._pwm = pwm


def size(pwm):
    return pwm._mat.shape[0]


def subPWM(i, l, pwm):
    """M.subMatrix (i,0) (i+l,3): rows i..i+l inclusive, as the reference passes them (Motif.hs:72-73)."""
    return PWM(pwm._nSites, pwm._mat[i:i + l + 1])


def rcPWM(pwm):
    """Reverse of the row-major flattening (Motif.hs:77-78)."""
    return PWM(pwm._nSites, pwm._mat.reshape(-1)[::-1].reshape(-1, 4).copy())


def gcContentPWM(pwm):
    acc = 0.0
    for i in range(size(pwm)):
        acc = acc + pwm._mat[i, 1] + pwm._mat[i, 2]
    return acc / size(pwm)


def _pow10(k):
    """`10 ^ k :: Double` as GHC evaluates it: exponentiation by repeated squaring IN Double (GHC.Real.(^): powImpl /
    powImplAcc), so it is exact up to 10^22 and carries the squaring's rounding errors beyond."""
    if k == 0:
        return 1.0
    x, y = 10.0, k
    while y % 2 == 0:
        x, y = x * x, y // 2
    if y == 1:
        return x
    z, x, y = x, x * x, y // 2
    while True:
        if y % 2 == 0:
            x, y = x * x, y // 2
        elif y == 1:
            return x * z
        else:
            z, x, y = x * z, x * x, y // 2


def read_double(tok):
    """Bio.Utils.Misc.readDouble (Utils/Misc.hs:21-25) = `fst . fromMaybe err . readSigned readExponential` of the
    third-party `bytestring-lexing` (>= 0.5, bioinformatics-toolkit.cabal:68; source not under /root/reference —
    restated from the published 0.5.0.x algorithm, Data.ByteString.Lex.Fractional):

      * optional sign, then DIGITS [ '.' DIGITS ] [ (e|E) [sign] DIGITS ]; at least one digit before the point; whatever
        follows the longest such prefix is ignored (`fst`);
      * the digits are accumulated EXACTLY as one Integer mantissa (whole * 10^len(frac) + frac), scale = -len(frac)
        plus the exponent;
      * value = 0 if the mantissa is 0, else `fromInteger mantissa / 10 ^ (-scale)` (scale < 0) or
        `fromInteger mantissa * 10 ^ scale`, with `10 ^ k` evaluated in Double by repeated squaring.

    So it is NOT correctly rounded like strtod / Python float(): a mantissa above 2^53 is rounded once by fromInteger
    and again by the division, and powers of ten beyond 10^22 are inexact.  Version assumption: fromInteger rounds to
    nearest-even (true for mantissas below 2^63; GHC's big-Integer conversion differs between GHC versions — unpinned).
    """
    b = tok if isinstance(tok, (bytes, bytearray)) else str(tok).encode("latin1")

    def fail():
        raise ValueError("readDouble: Fail to cast ByteString to Double:%r" % (bytes(b),))

    i, n, neg = 0, len(b), False
    if n and b[0] in b"+-":
        neg, i = b[0] == 0x2D, 1
    j = i
    while j < n and 0x30 <= b[j] <= 0x39:
        j += 1
    if j == i:
        fail()
    mant, scale = int(b[i:j]), 0
    if j < n and b[j] == 0x2E:
        k = j + 1
        while k < n and 0x30 <= b[k] <= 0x39:
            k += 1
        if k > j + 1:
            mant, scale, j = mant * 10 ** (k - j - 1) + int(b[j + 1:k]), -(k - j - 1), k
    if j < n and b[j] in b"eE":
        k = j + 1
        eneg = False
        if k < n and b[k] in b"+-":
            eneg, k = b[k] == 0x2D, k + 1
        k0 = k
        while k < n and 0x30 <= b[k] <= 0x39:
            k += 1
        if k > k0:
            scale += -int(b[k0:k]) if eneg else int(b[k0:k])
    if mant == 0:
        v = 0.0
    elif scale < 0:
        v = float(mant) / _pow10(-scale)
    else:
        v = float(mant) * _pow10(scale)
    return -v if neg else v


def toPWM(lines):
    """Motif.hs:329-330: lines of 4 numbers, each through readDouble (Utils/Misc.hs:21-25)."""
    return PWM(None, [[read_double(x) for x in (l if isinstance(l, (bytes, bytearray)) else l.encode("latin1")).split()] for l in lines])


def _argmax_last(row):
    bi, bs = 0, row[0]
    for i in range(1, 4):
        if not (row[i] < bs):
            bs, bi = row[i], i
    return bi


def optimalScore(bg, pwm):
    """Motif.hs:152-162 (libm log on the host, like GHC)."""
    f = bg.freqs
    acc = 0.0
    for row in pwm._mat:
        j = _argmax_last(row)
        acc = acc + math.log(row[j] / f[j])
    return acc


# ---- device handles ------------------------------------------------------------------------
def _dev_seq(dna, keep_ascii=False, ctx=None):
    if isinstance(dna, _lib.DeviceSeq):
        return dna
    ctx = ctx or _lib.default_context()
    bs = dna.bs if isinstance(dna, DNA) else bytes(dna)
    return _lib.DeviceSeq(ctx, bs, keep_ascii=keep_ascii)


def _dev_motifs(bg, pwms, ctx=None):
    ctx = ctx or _lib.default_context()
    return _lib.DeviceMotifs(ctx, [p._mat for p in pwms], bg.freqs)


def scores(bg, pwm, dna):
    """Scores of a long sequence at each position (Motif.hs:127-133), IUPAC aware."""
    seq = _dev_seq(dna, keep_ascii=True)
    return _lib.scores(seq, _dev_motifs(bg, [pwm], seq.ctx), 0, 0)


def scores_(bg, pwm, dna):
    """scores' — the streaming version (Motif.hs:136-145); here an iterator over the same array."""
    return iter(scores(bg, pwm, dna))


def score(bg, pwm, dna):
    """Motif.hs:147-149: scoreHelp over the whole of `dna` (its length should be size pwm)."""
    bs = dna.bs if isinstance(dna, DNA) else bytes(dna)
    n = size(pwm)
    if len(bs) == n:
        return float(scores(bg, pwm, DNA(bs))[0])
    # B.foldl over every byte with M.unsafeIndex: only len == n is meaningful in the reference
    raise ValueError("score: sequence length must equal the PWM size")


# ---- CDF ----------------------------------------------------------------------------------------
class CDF:
    """newtype CDF = CDF (Vector (x, P(X <= x))) (Motif.hs:233)."""

    def __init__(self, xs, ps):
        self.xs = np.ascontiguousarray(xs, dtype=np.float64)
        self.ps = np.ascontiguousarray(ps, dtype=np.float64)

    def __len__(self):
        return int(self.xs.size)


def scoreCDF(bg, pwm):
    """Motif.hs:282-326, computed on the device."""
    xs, ps = _dev_motifs(bg, [pwm]).score_cdf(0)
    return CDF(xs, ps)


def _lower_bound(keys, q):
    # binarySearchBy (Utils/Functions.hs:136-152) returns the first index whose key is >= q
    return int(np.searchsorted(keys, q, side="left"))


def cdf(c, x):
    """P(X <= x) (Motif.hs:236-249)."""
    n = len(c)
    i = _lower_bound(c.xs, x)
    if i >= n:
        return float(c.ps[-1])
    if i == 0:
        return float(c.ps[0]) if x == c.xs[0] else 0.0
    a, a1, b, b1 = float(c.xs[i - 1]), float(c.ps[i - 1]), float(c.xs[i]), float(c.ps[i])
    al = (b - x) / (b - a)
    be = (x - a) / (b - a)
    return al * a1 + be * b1


def cdf_(c, p):
    """cdf' — the inverse of cdf (Motif.hs:252-270)."""
    if p > 1 or p < 0:
        raise ValueError("p must be in [0,1]")
    n = len(c)
    i = _lower_bound(c.ps, p)
    if i >= n:
        return math.inf
    if i == 0:
        if p == c.ps[0]:
            return float(c.xs[0])
        raise ValueError("cdf': impossible!")
    a, a1, b, b1 = float(c.xs[i - 1]), float(c.ps[i - 1]), float(c.xs[i]), float(c.ps[i])
    al = (b1 - p) / (b1 - a1)
    be = (p - a1) / (b1 - a1)
    return al * a + be * b


def truncateCDF(x, c):
    """Motif.hs:273-274."""
    k = c.ps >= x
    return CDF(c.xs[k], c.ps[k])


def pValueToScore(p, bg, pwm):
    """Motif.hs:213-214: cdf' (scoreCDF bg pwm) (1 - p), evaluated on the device."""
    return float(_dev_motifs(bg, [pwm]).pvalue_to_score(p)[0])


def pValueToScoreExact(p, bg, pwm):
    """Motif.hs:218-230: the exact cutoff by enumerating all 4^l k-mers (`replicateM l "ACGT"`: first letter slowest),
    scoring each with scoreHelp (sequential fp64 sum of libm logs, pseudocount 0.0001, Motif.hs:164-193) and pBkgd
    (sequential product, 197-208), sorting by descending score and accumulating the background probability until it
    exceeds p.  The reference calls it "inpractical for long motifs" and uses it as a test oracle only; so does this
    host restatement (l <= 13).  Ties in the score keep enumeration order (the reference's introsort is unstable; the
    returned score does not depend on the order of ties)."""
    l = size(pwm)
    if l > 13:
        raise ValueError("pValueToScoreExact: 4^%d k-mers is impractical" % l)
    f = bg.freqs
    sc = np.zeros(1)
    pb = np.ones(1)
    for k in range(l):
        row = [0.0001 if v == 0 else float(v) for v in pwm._mat[k]]
        t = np.array([math.log(row[b] / f[b]) for b in range(4)])
        sc = (sc[:, None] + t[None, :]).reshape(-1)           # acc + sc, left to right; the new letter varies fastest
        pb = (pb[:, None] * np.array(f)[None, :]).reshape(-1)
    order = np.argsort(-sc, kind="stable")
    acc = np.cumsum(pb[order])                                # go: acc + snd, strictly sequential
    j = int(np.searchsorted(acc > p, True))
    if j >= acc.size:
        raise IndexError("pValueToScoreExact: index out of bounds (p >= total probability)")
    return float(sc[order[j]])


def pValueToScores(p, bg, pwms):
    """Batched pValueToScore for a list of PWMs (one launch per DP column for all motifs)."""
    return _dev_motifs(bg, pwms).pvalue_to_score(p)


def toIUPAC(pwm):
    """Consensus sequence (Motif.hs:107-124)."""
    out = []
    for row in pwm._mat:
        order = sorted(zip("ACGT", row), key=lambda t: -t[1])   # sortBy (flip (comparing snd)): stable
        (a, pa), (b, pb) = order[0], order[1]
        if pa > 0.5 and pa > 2 * pb:
            out.append(a)
        elif pa + pb > 0.75:
            x, y = (a, b) if a <= b else (b, a)
            out.append({("A", "C"): "M", ("G", "T"): "K", ("A", "T"): "W", ("C", "G"): "S", ("C", "T"): "Y", ("A", "G"): "R"}[(x, y)])
        else:
            out.append("N")
    return DNA("".join(out).encode(), "IUPAC")


# ---- motif files -----------------------------------------------------------------------------
def fromMEME(meme):
    """Motif.hs:361-386 (state machine over the lines of a MEME file)."""
    if isinstance(meme, bytes):
        meme = meme.decode("latin1")
    out, st, cur = [], 0, None
    for x in meme.split("\n"):
        if x.startswith("MOTIF"):
            st, cur = 1, [x[6:]]
            continue
        if st == 1:
            if x.startswith("letter-probability matrix:"):
                cur.append(x.split()[7])
                st = 2
        elif st == 2:
            y = x.lstrip(" ")
            if y == "" or y.startswith("URL"):
                out.append(_to_motif(cur))
                st, cur = 0, None
            else:
                cur.append(y)
    if st == 2:
        out.append(_to_motif(cur))
    return out


def _to_motif(parts):
    name, n, rows = parts[0], parts[1], parts[2:]
    return Motif(name.encode("latin1"), PWM(int(n), [[read_double(v) for v in r.split()] for r in rows]))


def readMEME(path):
    with open(path, "rb") as f:
        return fromMEME(f.read())


def _shortest(x):
    """Data.Double.Conversion.ByteString.toShortest (the `double-conversion` library's ToShortest, i.e. the ECMAScript
    Number-to-string rules): shortest round-trip digits; plain decimals for 1e-6 <= |x| < 1e21 ("0.00001", "1", "123.5"),
    otherwise d[.ddd]e[+-]n without zero padding ("1e-7", "1.5e+21").  Used by toMEME / writeFasta (Motif.hs:334,355)
    and `merge_motifs dist` (MergeMotifs.hs:167)."""
    x = float(x)
    r = repr(x)
    if "e" not in r and "n" not in r:                # positional repr (1e-4 <= |x| < 1e16): already the ECMAScript form
        return r[:-2] if r.endswith(".0") else r     # ... except for the trailing ".0" (and "-0.0", never produced here)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "0"                                   # ToShortest prints -0.0 as "-0"; probabilities / distances are never -0
    sign = "-" if x < 0 else ""
    r = repr(abs(x))                                 # shortest round-trip digits (David Gay / Ryu-equivalent)
    mant, _, ex = r.partition("e")
    ip, _, fp = mant.partition(".")
    if fp == "0":
        fp = ""
    digits = (ip + fp).lstrip("0")
    n = len(ip) + (int(ex) if ex else 0)             # value = 0.DIGITS x 10^n, with leading zeros of ip+fp accounted for
    n -= len(ip + fp) - len((ip + fp).lstrip("0"))
    digits = digits.rstrip("0") or "0"
    k = len(digits)
    if k <= n <= 21:
        return sign + digits + "0" * (n - k)
    if 0 < n <= 21:
        return sign + digits[:n] + "." + digits[n:]
    if -6 < n <= 0:
        return sign + "0." + "0" * (-n) + digits
    e = n - 1
    body = digits if k == 1 else digits[0] + "." + digits[1:]
    return sign + body + "e" + ("+" if e >= 0 else "-") + str(abs(e))


def _show_ffloat(x):
    """Text.Printf's `%f` WITHOUT a precision = `showFFloat Nothing` (Numeric): all shortest round-trip digits in fixed
    notation, never an exponent, at least one digit on either side of the point ("0.25", "1.0", "0.0000001") — not C's
    six decimals.  Used by toMEME's background line (Motif.hs:351)."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    sign = "-" if (x < 0 or (x == 0 and math.copysign(1.0, x) < 0)) else ""
    mant, _, ex = repr(abs(x)).partition("e")
    ip, _, fp = mant.partition(".")
    digits = (ip + fp).lstrip("0")
    e = len(ip) + (int(ex) if ex else 0) - (len(ip + fp) - len(digits))      # value = 0.DIGITS x 10^e (floatToDigits)
    digits = digits.rstrip("0")
    if not digits:
        return sign + "0.0"
    if e <= 0:
        return sign + "0." + "0" * (-e) + digits
    whole, rest = digits[:e].ljust(e, "0"), digits[e:]
    return sign + whole + "." + (rest or "0")


def toMEME(motifs, bg=default_bkgd):
    a, c, g, t = bg.freqs
    s = "MEME version 4\n\nALPHABET= ACGT\n\nstrands: + -\n\nBackground letter frequencies\nA %s C %s G %s T %s\n\n" % tuple(_show_ffloat(v) for v in (a, c, g, t))
    for m in motifs:
        sites = 1 if m._pwm._nSites is None else m._pwm._nSites
        s += "MOTIF " + m._name.decode("latin1") + "\n"
        s += "letter-probability matrix: alength= 4 w= %d nsites= %d E= 0\n" % (size(m._pwm), sites)
        s += "".join(" " + " ".join(_shortest(v) for v in row) + "\n" for row in m._pwm._mat) + "\n"
    return s.encode("latin1")


def writeMEME(path, motifs, bg=default_bkgd):
    with open(path, "wb") as f:
        f.write(toMEME(motifs, bg))


def readFasta_(path):
    """readFasta' for Motif records (Data/Fasta.hs:33-51): '>name' then rows of 4 numbers."""
    out, cur = [], None
    with open(path, "rb") as f:
        for l in f.read().split(b"\n"):
            if l == b"":
                continue
            if l[:1] == b">":
                if cur is not None:
                    out.append(Motif(cur[0], toPWM(cur[1])))
                cur = (l[1:], [])
            elif cur is not None:
                cur[1].append(l)
    if cur is not None:
        out.append(Motif(cur[0], toPWM(cur[1])))
    return out


def writeFasta(path, motifs):
    with open(path, "wb") as f:
        for m in motifs:
            f.write(b">" + m._name + b"\n")
            f.write("".join(" ".join(_shortest(v) for v in row) + "\n" for row in m._pwm._mat).encode() + b"\n")
