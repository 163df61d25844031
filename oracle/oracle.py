"""ctypes wrapper over oracle/liboracle.so — the CPU restatement of the reference's algorithm.

TEST INFRASTRUCTURE ONLY (see oracle/pwm_oracle.cpp header): imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs, never by the
product package.  PARITY STATUS: parity unpinned for absolute values (no GHC here); pinned only
by the reference's five test invariants (bioinformatics-toolkit/tests/Tests/Motif.hs:58-97).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_optimal_score.restype = C.c_double
        L.orc_max_match_sc.restype = C.c_double
        L.orc_cdf.restype = C.c_double
        L.orc_cdf_inv.restype = C.c_double
        L.orc_pvalue_to_score.restype = C.c_double
        L.orc_pvalue_to_score_exact.restype = C.c_double
        for f in ("orc_find_tfbs", "orc_scores", "orc_find_tfbs_slow", "orc_score_cdf", "orc_truncate_cdf",
                  "orc_to_affinity", "orc_scan_mt"):
            getattr(L, f).restype = C.c_int64
        _LIB = L
    return _LIB


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t):
    return a.ctypes.data_as(t)


def _seq(dna):
    if isinstance(dna, (bytes, bytearray)):
        dna = np.frombuffer(bytes(dna), dtype=np.uint8)
    return np.ascontiguousarray(dna, dtype=np.uint8)


DEF_BG = (0.25, 0.25, 0.25, 0.25)


def rc_pwm(mat):
    mat = _d(mat)
    out = np.empty_like(mat)
    lib().orc_rc_pwm(C.c_int(mat.shape[0]), _p(mat, _dp), _p(out, _dp))
    return out


def optimal_scores_suffix(bg, mat):
    mat, bg = _d(mat), _d(bg)
    out = np.empty(mat.shape[0])
    lib().orc_optimal_scores_suffix(_p(bg, _dp), C.c_int(mat.shape[0]), _p(mat, _dp), _p(out, _dp))
    return out


def optimal_score(bg, mat):
    mat, bg = _d(mat), _d(bg)
    return lib().orc_optimal_score(_p(bg, _dp), C.c_int(mat.shape[0]), _p(mat, _dp))


def score_table(bg, mat):
    mat, bg = _d(mat), _d(bg)
    out = np.empty_like(mat)
    lib().orc_score_table(_p(bg, _dp), C.c_int(mat.shape[0]), _p(mat, _dp), _p(out, _dp))
    return out


def find_tfbs(bg, mat, dna, thres, skip=True, sigma=None, mode=0, cap=None):
    mat, bg, dna = _d(mat), _d(bg), _seq(dna)
    cap = int(cap if cap is not None else max(16, dna.size))
    pos = np.empty(cap, dtype=np.int64)
    sc = np.empty(cap, dtype=np.float64)
    sg = _p(_d(sigma), _dp) if sigma is not None else None
    n = lib().orc_find_tfbs(_p(bg, _dp), C.c_int(mat.shape[0]), _p(mat, _dp), sg, _p(dna, _u8p), C.c_int64(dna.size),
                            C.c_double(thres), C.c_int(1 if skip else 0), C.c_int(mode), _p(pos, _i64p), _p(sc, _dp),
                            C.c_int64(cap))
    if n < 0:
        raise ValueError("Bio.Motif.Search.lookAheadSearch: invalid nucleotide")
    if n > cap:
        return find_tfbs(bg, mat, dna, thres, skip, sigma, mode, cap=n)
    return pos[:n].copy(), sc[:n].copy()


def find_tfbs_slow(bg, mat, dna, thres):
    mat, bg, dna = _d(mat), _d(bg), _seq(dna)
    cap = max(16, dna.size)
    pos = np.empty(cap, dtype=np.int64)
    sc = np.empty(cap, dtype=np.float64)
    n = lib().orc_find_tfbs_slow(_p(bg, _dp), C.c_int(mat.shape[0]), _p(mat, _dp), _p(dna, _u8p), C.c_int64(dna.size),
                                 C.c_double(thres), _p(pos, _i64p), _p(sc, _dp), C.c_int64(cap))
    if n < 0:
        raise ValueError("Bio.Motif.score: invalid nucleotide")
    return pos[:n].copy(), sc[:n].copy()


def scores(bg, mat, dna):
    mat, bg, dna = _d(mat), _d(bg), _seq(dna)
    out = np.empty(max(0, dna.size - mat.shape[0] + 1))
    n = lib().orc_scores(_p(bg, _dp), C.c_int(mat.shape[0]), _p(mat, _dp), _p(dna, _u8p), C.c_int64(dna.size), _p(out, _dp))
    if n < 0:
        raise ValueError("Bio.Motif.score: invalid nucleotide")
    return out


def max_match_sc(bg, mat, dna):
    mat, bg, dna = _d(mat), _d(bg), _seq(dna)
    ok = C.c_int(0)
    v = lib().orc_max_match_sc(_p(bg, _dp), C.c_int(mat.shape[0]), _p(mat, _dp), _p(dna, _u8p), C.c_int64(dna.size), C.byref(ok))
    if not ok.value:
        raise ValueError("Bio.Motif.Search.lookAheadSearch: invalid nucleotide")
    return v


def score_cdf(bg, mat):
    mat, bg = _d(mat), _d(bg)
    cap = 200001
    xs = np.empty(cap)
    ps = np.empty(cap)
    n = lib().orc_score_cdf(_p(bg, _dp), C.c_int(mat.shape[0]), _p(mat, _dp), _p(xs, _dp), _p(ps, _dp), C.c_int64(cap))
    return xs[:n].copy(), ps[:n].copy()


def cdf(cdf_pts, x):
    xs, ps = _d(cdf_pts[0]), _d(cdf_pts[1])
    return lib().orc_cdf(_p(xs, _dp), _p(ps, _dp), C.c_int64(xs.size), C.c_double(x))


def cdf_inv(cdf_pts, p):
    xs, ps = _d(cdf_pts[0]), _d(cdf_pts[1])
    st = C.c_int(0)
    v = lib().orc_cdf_inv(_p(xs, _dp), _p(ps, _dp), C.c_int64(xs.size), C.c_double(p), C.byref(st))
    if st.value == 1:
        raise ValueError("p must be in [0,1]")
    if st.value == 2:
        raise ValueError("cdf': impossible!")
    return v


def truncate_cdf(x, cdf_pts):
    xs, ps = _d(cdf_pts[0]), _d(cdf_pts[1])
    ox, op = np.empty_like(xs), np.empty_like(ps)
    n = lib().orc_truncate_cdf(_p(xs, _dp), _p(ps, _dp), C.c_int64(xs.size), C.c_double(x), _p(ox, _dp), _p(op, _dp))
    return ox[:n].copy(), op[:n].copy()


def pvalue_to_score(p, bg, mat):
    mat, bg = _d(mat), _d(bg)
    st = C.c_int(0)
    v = lib().orc_pvalue_to_score(C.c_double(p), _p(bg, _dp), C.c_int(mat.shape[0]), _p(mat, _dp), C.byref(st))
    if st.value:
        raise ValueError("cdf' failed: status %d" % st.value)
    return v


def pvalue_to_score_exact(p, bg, mat):
    mat, bg = _d(mat), _d(bg)
    return lib().orc_pvalue_to_score_exact(C.c_double(p), _p(bg, _dp), C.c_int(mat.shape[0]), _p(mat, _dp))


def to_affinity(x):
    return int(lib().orc_to_affinity(C.c_double(x)))


PENAL = {"linear": 0, "quadratic": 1, "cubic": 2, "exp": 3}
COMB = {"l1": 0, "l2": 1, "l3": 2, "max": 3}


def alignment(m1, m2, penal="exp", gap=0.05, comb="l1"):
    m1, m2 = _d(m1), _d(m2)
    d = C.c_double(0)
    sd = C.c_int(0)
    pos = C.c_int(0)
    lib().orc_alignment(C.c_int(m1.shape[0]), _p(m1, _dp), C.c_int(m2.shape[0]), _p(m2, _dp), C.c_int(PENAL[penal]),
                        C.c_double(gap), C.c_int(COMB[comb]), C.byref(d), C.byref(sd), C.byref(pos))
    return d.value, (bool(sd.value), pos.value)


def _pack_mats(mats):
    ncol = np.array([m.shape[0] for m in mats], dtype=np.int32)
    off = np.zeros(len(mats), dtype=np.int64)
    if len(mats) > 1:
        off[1:] = np.cumsum(ncol[:-1].astype(np.int64) * 4)
    flat = np.concatenate([_d(m).reshape(-1) for m in mats]) if mats else np.zeros(0)
    return ncol, off, flat


def alignment_all_pairs(mats, penal="exp", gap=0.05, comb="l1", nthreads=1):
    M = len(mats)
    ncol, off, flat = _pack_mats(mats)
    npair = M * (M - 1) // 2
    d = np.empty(npair)
    sd = np.empty(npair, dtype=np.uint8)
    pos = np.empty(npair, dtype=np.int32)
    lib().orc_alignment_all_pairs(C.c_int(M), _p(ncol, _i32p), _p(flat, _dp), _p(off, _i64p), C.c_int(PENAL[penal]),
                                  C.c_double(gap), C.c_int(COMB[comb]), C.c_int(nthreads), _p(d, _dp), _p(sd, _u8p),
                                  _p(pos, _i32p))
    return d, sd, pos


def scan_mt(bg, mats, thres, dna, mode=1, nthreads=1, chunk=500000, cap=0):
    """Both-strand thresholded scan of all motifs (CPU baseline driver). Returns (count, hits or None)."""
    bg, dna = _d(bg), _seq(dna)
    ncol, off, flat = _pack_mats(mats)
    thres = _d(thres)
    cap = int(cap)
    om = np.empty(max(cap, 1), dtype=np.int32)
    os_ = np.empty(max(cap, 1), dtype=np.uint8)
    op = np.empty(max(cap, 1), dtype=np.int64)
    osc = np.empty(max(cap, 1), dtype=np.float64)
    n = lib().orc_scan_mt(_p(bg, _dp), C.c_int(len(mats)), _p(ncol, _i32p), _p(flat, _dp), _p(off, _i64p), _p(thres, _dp),
                          _p(dna, _u8p), C.c_int64(dna.size), C.c_int(mode), C.c_int(nthreads), C.c_int64(chunk),
                          _p(om, _i32p), _p(os_, _u8p), _p(op, _i64p), _p(osc, _dp), C.c_int64(cap))
    if cap == 0:
        return int(n), None
    k = min(int(n), cap)
    order = np.lexsort((op[:k], os_[:k], om[:k]))
    return int(n), (om[:k][order], os_[:k][order], op[:k][order], osc[:k][order])
