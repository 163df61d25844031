"""ctypes binding of libpwmscan.so (include/pwmscan.h) — the only compute path of this package.

There is no CPU fallback: if the shared library is missing or no CUDA device is present every
compute entry point raises `PwmScanError` (loudly), it never silently degrades.
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpwmscan.so")

SEQ_UPPERCASE = 1
SEQ_KEEP_ASCII = 2

OK, EINVAL, ECUDA, ENUCL, ERANGE, EDOMAIN = 0, -1, -2, -3, -4, -5

PENAL = {"linear": 0, "quadratic": 1, "cubic": 2, "exp": 3}
COMB = {"l1": 0, "l2": 1, "l3": 2, "max": 3}


class PwmScanError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pwmscan error %d: %s" % (code, msg))
        self.code = code
        self.msg = msg


class InvalidNucleotide(PwmScanError, ValueError):
    """The reference's `error "... invalid nucleotide"` (Motif/Search.hs:157, Motif.hs:185)."""


_lib = None
_lib_lock = threading.Lock()

_vp = C.c_void_p
_dp = C.POINTER(C.c_double)


def build(verbose=False):
    """Compile libpwmscan.so in-tree with nvcc for sm_100a (see csrc/Makefile)."""
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc")] + ([] if verbose else ["-s"]))


def load():
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise PwmScanError(ECUDA, "libpwmscan.so is not built (%s); run bioinformatics_toolkit_b200._lib.build() — "
                                      "there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.pwmscan_last_error.restype = C.c_char_p
        L.pwmscan_last_error.argtypes = [_vp]
        L.pwmscan_seq_length.restype = C.c_int64
        L.pwmscan_packed_units.restype = C.c_int64
        L.pwmscan_packed_units.argtypes = [C.c_int64]
        L.pwmscan_hits_count.restype = C.c_int64
        L.pwmscan_hits_count.argtypes = [_vp]
        L.pwmscan_hits_first_invalid.argtypes = [_vp, C.c_int, C.POINTER(C.c_int64)]
        L.pwmscan_ctx_destroy.argtypes = [_vp]
        L.pwmscan_seq_free.argtypes = [_vp]
        L.pwmscan_motifs_free.argtypes = [_vp]
        L.pwmscan_plan_free.argtypes = [_vp]
        L.pwmscan_hits_free.argtypes = [_vp]
        L.pwmscan_free.argtypes = [_vp]
        L.pwmscan_alnset_free.argtypes = [_vp]
        _lib = L
        return L


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a):
    return a.ctypes.data_as(_vp)


def _alphabet_code(alphabet):
    """0 = `DNA Basic` (ACGT), 1 = `DNA IUPAC` (Seq.hs:41-57); anything else is a caller bug."""
    try:
        return {"Basic": 0, "IUPAC": 1}[alphabet]
    except KeyError:
        raise ValueError("alphabet must be 'Basic' or 'IUPAC', got %r" % (alphabet,))


class Context:
    """One per device (pwmscan_ctx)."""

    def __init__(self, device=0):
        self.L = load()
        self.h = _vp()
        rc = self.L.pwmscan_ctx_create(C.c_int(device), C.byref(self.h))
        if rc != 0:
            raise PwmScanError(rc, (self.L.pwmscan_last_error(None) or b"").decode())
        self.device = device
        # Handles made from this context (sequences, motif sets, plans, hit lists, alignment sets) hand buffers back to ITS pools
        # when they are freed, so the C context must outlive them all.  Reference counts alone do not guarantee that: the
        # cyclic garbage collector (module globals captured by lambdas, interpreter shutdown) runs the finalisers of a
        # cycle in arbitrary order.  The context therefore counts its live handles and is destroyed by whoever goes last.
        self._live = 0
        self._closing = False

    def _retain(self):
        self._live += 1

    def _release(self):
        self._live -= 1
        if self._closing and self._live <= 0:
            self._destroy()

    def _destroy(self):
        if self.h:
            self.L.pwmscan_ctx_destroy(self.h)
            self.h = _vp()

    def check(self, rc):
        if rc == 0:
            return
        msg = (self.L.pwmscan_last_error(self.h) or b"").decode()
        if rc == ENUCL:
            raise InvalidNucleotide(rc, msg)
        raise PwmScanError(rc, msg)

    def timing(self):
        """Device timing of the last scan call on this context (summed over the items of a scan_many)."""
        out = np.zeros(12)
        self.L.pwmscan_last_timing(self.h, _ptr(out), C.c_int(12))
        return {"prefilter_ms": out[0], "rescore_ms": out[1], "order_ms": out[2], "device_ms": out[3],
                "h2d_bytes": out[4], "d2h_bytes": out[5], "candidates": out[6], "prefilter_launches": out[7],
                "kernel_launches": out[8], "hits": out[9],
                # the auxiliary entry points (scoreCDF, alignment, dense scores, per-hit affinity) report here and leave the scan slots alone
                "aux_kernel_ms": out[10], "aux_host_fallbacks": out[11]}

    def scan_host_first_invalid(self, alphabet="IUPAC"):
        """fromBS verdict of the bytes of the last Plan.scan_host on this context: offset of the first byte
        outside the alphabet ("Basic" = ACGT, "IUPAC"; the names of seq.ALPHABET), or -1."""
        pos = C.c_int64(-1)
        self.check(self.L.pwmscan_scan_host_first_invalid(self.h, C.c_int(_alphabet_code(alphabet)), C.byref(pos)))
        return int(pos.value)

    def close(self):
        """Destroys the context — at once if nothing made from it is alive, else when its last handle is freed."""
        self._closing = True
        if self._live <= 0:
            self._destroy()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=None):
    """Process-wide context for `device` (default: LOCAL_RANK or 0)."""
    if device is None:
        device = int(os.environ.get("PWMSCAN_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class DeviceSeq:
    """A DNA sequence (or several segments) packed in HBM (pwmscan_seq)."""

    def __init__(self, ctx, data, seg_off=None, uppercase=False, keep_ascii=False):
        self.ctx = ctx
        if isinstance(data, (bytes, bytearray, memoryview)):
            data = np.frombuffer(bytes(data), dtype=np.uint8)
        data = _arr(data, np.uint8)
        flags = (SEQ_UPPERCASE if uppercase else 0) | (SEQ_KEEP_ASCII if keep_ascii else 0)
        self.h = _vp()
        self.length = int(data.size)
        self.keep_ascii = keep_ascii
        if seg_off is None:
            self.nseg = 1
            self.seg_off = np.array([0, data.size], dtype=np.int64)
            rc = ctx.L.pwmscan_seq_upload(ctx.h, _ptr(data), C.c_int64(data.size), C.c_int(flags), C.byref(self.h))
        else:
            so = _arr(seg_off, np.int64)
            self.nseg = int(so.size - 1)
            self.seg_off = so
            rc = ctx.L.pwmscan_seq_upload_segments(ctx.h, _ptr(data), _ptr(so), C.c_int32(self.nseg), C.c_int(flags), C.byref(self.h))
        ctx.check(rc)
        ctx._retain()

    @classmethod
    def from_packed(cls, ctx, packed, start=0, end=None):
        """pwmscan_seq_upload_packed: bases [start, end) of a PackedSeq (start a multiple of 32) -> a resident sequence."""
        seq2, mask, n = packed.slice(start, end)
        self = cls.__new__(cls)
        self.ctx = ctx
        self.h = _vp()
        self.length = n
        self.keep_ascii = False
        self.nseg = 1
        self.seg_off = np.array([0, n], dtype=np.int64)
        ctx.check(ctx.L.pwmscan_seq_upload_packed(ctx.h, _ptr(seq2), _ptr(mask), C.c_int64(n), C.byref(self.h)))
        ctx._retain()
        return self

    def first_invalid(self, alphabet="IUPAC"):
        pos = C.c_int64(-1)
        self.ctx.check(self.ctx.L.pwmscan_seq_first_invalid(self.ctx.h, self.h, C.c_int(_alphabet_code(alphabet)), C.byref(pos)))
        return pos.value

    def free(self):
        if self.h:
            self.ctx.L.pwmscan_seq_free(self.h)
            self.h = _vp()
            self.ctx._release()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PackedPlanes:
    """The two planes of a packed sequence, wherever they live (arrays in memory, a memory-mapped `.packed` sidecar of the
    genome index): `seq2` uint32[2 * units], `mask` uint32[units], units = ceil(length / 32)."""

    def __init__(self, seq2, mask, length, first_non_acgt=-1, first_non_iupac=-1):
        self.seq2, self.mask, self.length = seq2, mask, int(length)
        self.first_non_acgt, self.first_non_iupac = int(first_non_acgt), int(first_non_iupac)

    def slice(self, start=0, end=None):
        """(seq2 view, mask view, length) of bases [start, end): start must be a multiple of 32."""
        end = self.length if end is None else min(int(end), self.length)
        if start % 32 or not (0 <= start <= end):
            raise ValueError("a packed slice starts at a multiple of 32 bases")
        return self.seq2[start // 16:], self.mask[start // 32:], end - start


class PackedSeq(PackedPlanes):
    """The packed genome store of one sequence in HOST memory (pwmscan_pack_host): 2 bits per base + 1 invalid-base bit per
    base, 0.375 bytes per base, made once (mkIndex time) and uploaded instead of the bytes.  Pure host code — works
    without a GPU."""

    def __init__(self, data, uppercase=False, nthreads=0, pinned=False):
        L = load()
        if isinstance(data, (bytes, bytearray, memoryview)):
            data = np.frombuffer(bytes(data), dtype=np.uint8)
        data = _arr(data, np.uint8)
        self.length = int(data.size)
        units = (self.length + 31) // 32
        if pinned:
            import torch
            self._keep = (torch.empty(2 * units, dtype=torch.int32).pin_memory(), torch.empty(units, dtype=torch.int32).pin_memory())
            self.seq2 = self._keep[0].numpy().view(np.uint32)
            self.mask = self._keep[1].numpy().view(np.uint32)
        else:
            self.seq2 = np.empty(2 * units, dtype=np.uint32)
            self.mask = np.empty(units, dtype=np.uint32)
        a, i = C.c_int64(-1), C.c_int64(-1)
        rc = L.pwmscan_pack_host(_ptr(data), C.c_int64(self.length), C.c_int(SEQ_UPPERCASE if uppercase else 0), _ptr(self.seq2), _ptr(self.mask),
                                 C.byref(a), C.byref(i), C.c_int(nthreads))
        if rc != OK:
            raise PwmScanError(rc, "pwmscan_pack_host: bad argument")
        self.first_non_acgt, self.first_non_iupac = a.value, i.value


class DeviceMotifs:
    """A set of PWMs + background with their device tables (pwmscan_motifs)."""

    def __init__(self, ctx, mats, bg=(0.25, 0.25, 0.25, 0.25)):
        self.ctx = ctx
        self.mats = [_arr(m, np.float64).reshape(-1, 4) for m in mats]
        self.M = len(self.mats)
        self.ncol = np.array([m.shape[0] for m in self.mats], dtype=np.int32)
        flat = np.concatenate([m.reshape(-1) for m in self.mats]) if self.mats else np.zeros(0)
        self.bg = _arr(bg, np.float64)
        self.h = _vp()
        ctx.check(ctx.L.pwmscan_motifs_create(ctx.h, C.c_int32(self.M), _ptr(self.ncol), _ptr(flat), _ptr(self.bg), C.byref(self.h)))
        ctx._retain()

    def set_sigma(self, m, strand, sigma):
        s = _arr(sigma, np.float64)
        if s.size != self.ncol[m]:
            raise ValueError("sigma must have one entry per PWM column")
        self.ctx.check(self.ctx.L.pwmscan_motifs_set_sigma(self.h, C.c_int32(m), C.c_int(strand), _ptr(s)))

    def get_sigma(self, m, strand=0):
        out = np.empty(int(self.ncol[m]))
        self.ctx.check(self.ctx.L.pwmscan_motifs_get_sigma(self.h, C.c_int32(m), C.c_int(strand), _ptr(out)))
        return out

    # -- scoreCDF / pValueToScore -------------------------------------------------------------
    def score_cdf(self, m):
        x = _dp()
        p = _dp()
        n = C.c_int64(0)
        self.ctx.check(self.ctx.L.pwmscan_score_cdf(self.ctx.h, self.h, C.c_int32(m), C.byref(x), C.byref(p), C.byref(n)))
        xs = np.ctypeslib.as_array(x, shape=(n.value,)).copy()
        ps = np.ctypeslib.as_array(p, shape=(n.value,)).copy()
        self.ctx.L.pwmscan_free(C.cast(x, _vp))
        self.ctx.L.pwmscan_free(C.cast(p, _vp))
        return xs, ps

    def pvalue_to_score(self, p):
        out = np.empty(self.M)
        self.ctx.check(self.ctx.L.pwmscan_pvalue_to_score(self.ctx.h, self.h, C.c_double(p), _ptr(out)))
        return out

    def cutoff_cdfs(self, p):
        """mkCutoffMotif for all motifs: (cutoffs[M], [(xs, ps) of truncateCDF (1 - 10p) cdf_m])."""
        th = np.empty(self.M)
        off = np.zeros(self.M + 1, dtype=np.int64)
        x = _dp()
        P = _dp()
        self.ctx.check(self.ctx.L.pwmscan_cutoff_cdfs(self.ctx.h, self.h, C.c_double(p), _ptr(th), C.byref(x), C.byref(P), _ptr(off)))
        tot = int(off[-1])
        xs = np.ctypeslib.as_array(x, shape=(max(tot, 1),))[:tot].copy()
        ps = np.ctypeslib.as_array(P, shape=(max(tot, 1),))[:tot].copy()
        self.ctx.L.pwmscan_free(C.cast(x, _vp))
        self.ctx.L.pwmscan_free(C.cast(P, _vp))
        return th, [(xs[off[m]:off[m + 1]], ps[off[m]:off[m + 1]]) for m in range(self.M)]

    # -- alignment ------------------------------------------------------------------------------
    def align_all_pairs(self, penal="exp", gap=0.05, comb="l1"):
        npair = self.M * (self.M - 1) // 2
        d = np.empty(npair)
        sd = np.empty(npair, dtype=np.uint8)
        pos = np.empty(npair, dtype=np.int32)
        self.ctx.check(self.ctx.L.pwmscan_align_all_pairs(self.ctx.h, self.h, C.c_int(PENAL[penal]), C.c_double(gap),
                                                          C.c_int(COMB[comb]), _ptr(d), _ptr(sd), _ptr(pos)))
        return d, sd, pos

    def align_pairs(self, a, b, penal="exp", gap=0.05, comb="l1"):
        a = _arr(a, np.int32)
        b = _arr(b, np.int32)
        d = np.empty(a.size)
        sd = np.empty(a.size, dtype=np.uint8)
        pos = np.empty(a.size, dtype=np.int32)
        self.ctx.check(self.ctx.L.pwmscan_align_pairs(self.ctx.h, self.h, C.c_int64(a.size), _ptr(a), _ptr(b), C.c_int(PENAL[penal]),
                                                      C.c_double(gap), C.c_int(COMB[comb]), _ptr(d), _ptr(sd), _ptr(pos)))
        return d, sd, pos

    def free(self):
        if self.h:
            self.ctx.L.pwmscan_motifs_free(self.h)
            self.h = _vp()
            self.ctx._release()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def format_dist(names, dist, a=None, b=None):
    """`merge_motifs dist` text through the library's host formatter (pwmscan_format_dist): one line
    name_a<TAB>name_b<TAB>toShortest d per pair; a = b = None means all i < j in packed order."""
    L = load()
    off = np.concatenate([[0], np.cumsum([len(x) for x in names])]).astype(np.int64)
    buf = b"".join(names)
    dist = _arr(dist, np.float64)
    text, tlen = C.c_char_p(), C.c_int64(0)
    if a is None:
        n, pa, pb = len(names) * (len(names) - 1) // 2, None, None
    else:
        a, b = _arr(a, np.int32), _arr(b, np.int32)
        n, pa, pb = a.size, _ptr(a), _ptr(b)
    if dist.size != n:
        raise ValueError("one distance per pair expected")
    rc = L.pwmscan_format_dist(C.c_int64(n), C.c_int32(len(names)), pa, pb, _ptr(dist), C.c_char_p(buf), _ptr(off),
                               C.byref(text), C.byref(tlen))
    if rc != 0:
        raise PwmScanError(rc, "pwmscan_format_dist failed")
    out = C.string_at(text, tlen.value)
    L.pwmscan_free.argtypes = [_vp]
    L.pwmscan_free(text)
    return out


class AlignSet:
    """PWMs resident on the device for iterativeMerge (pwmscan_alnset_*): all-pairs initialisation, per-merge slot
    update and row refresh without re-uploading the set."""

    def __init__(self, ctx, mats, penal="exp", gap=0.05, comb="l1"):
        self.ctx = ctx
        mats = [_arr(m, np.float64).reshape(-1, 4) for m in mats]
        self.M = len(mats)
        ncol = np.array([m.shape[0] for m in mats], dtype=np.int32)
        flat = np.concatenate([m.reshape(-1) for m in mats]) if mats else np.zeros(0)
        self.h = _vp()
        ctx.check(ctx.L.pwmscan_alnset_create(ctx.h, C.c_int32(self.M), _ptr(ncol), _ptr(flat), C.c_int(PENAL[penal]),
                                              C.c_double(gap), C.c_int(COMB[comb]), C.byref(self.h)))
        ctx._retain()

    def update(self, slot, mat):
        m = _arr(mat, np.float64).reshape(-1, 4)
        self.ctx.check(self.ctx.L.pwmscan_alnset_update(self.h, C.c_int32(slot), C.c_int32(m.shape[0]), _ptr(m)))

    def pairs(self, a=None, b=None):
        """align pwm_a[k] pwm_b[k]; a = b = None: all i < j (packed upper triangle, row-major)."""
        if a is None:
            n = self.M * (self.M - 1) // 2
            pa = pb = None
        else:
            a, b = _arr(a, np.int32), _arr(b, np.int32)
            n = a.size
            pa, pb = _ptr(a), _ptr(b)
        d = np.empty(n)
        sd = np.empty(n, dtype=np.uint8)
        pos = np.empty(n, dtype=np.int32)
        if n:
            self.ctx.check(self.ctx.L.pwmscan_alnset_pairs(self.h, C.c_int64(n), pa, pb, _ptr(d), _ptr(sd), _ptr(pos)))
        return d, sd, pos

    def _next(self, call):
        i, j, d, sd, pos = C.c_int32(-1), C.c_int32(-1), C.c_double(0.0), C.c_uint8(0), C.c_int32(0)
        self.ctx.check(call(C.byref(i), C.byref(j), C.byref(d), C.byref(sd), C.byref(pos)))
        return (None if i.value < 0 else (int(i.value), int(j.value), float(d.value), bool(sd.value), int(pos.value)))

    def matrix_init(self):
        """All-pairs matrix of iterativeMerge kept on the device; returns its first minimum (i, j, d, isSame, pos) or None."""
        return self._next(lambda *o: self.ctx.L.pwmscan_alnset_matrix_init(self.h, *o))

    def merge_step(self, i, j, mat):
        """Slot j has been merged into slot i (new PWM `mat`): update the resident matrix, return the next minimum or None."""
        m = _arr(mat, np.float64).reshape(-1, 4)
        return self._next(lambda *o: self.ctx.L.pwmscan_alnset_merge_step(self.h, C.c_int32(i), C.c_int32(j), C.c_int32(m.shape[0]), _ptr(m), *o))

    def free(self):
        if self.h:
            self.ctx.L.pwmscan_alnset_free(self.h)
            self.h = _vp()
            self.ctx._release()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _HitsHandle:
    """Owns a pwmscan_hits (pinned host columns); released when the last array view dies."""

    def __init__(self, ctx, h):
        self.ctx, self.L, self.h = ctx, ctx.L, h        # the hit buffer goes back to the context's pool: keep the context alive
        self.M = 0
        ctx._retain()

    def __del__(self):
        try:
            if self.h:
                self.L.pwmscan_hits_free(self.h)
                self.h = None
                self.ctx._release()
        except Exception:
            pass


class _OwnedArray(np.ndarray):
    """ndarray view over library-owned memory that keeps its owner alive (also through slices)."""

    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)


def _view(ptr, n, owner):
    a = np.ctypeslib.as_array(ptr, shape=(n,)).view(_OwnedArray)
    a._owner = owner
    return a


class Hits:
    """A hit list sorted by (segment, motif, strand, pos).

    The library transfers compact columns (12 bytes per hit for a single-segment scan): `score`, `pos32` and either
    `run_off` (single segment: hits of (motif m, strand s) are [run_off[2m+s], run_off[2m+s+1])) or `seg32` +
    `ms16` (= 2 * motif + strand).  `motif`, `strand`, `segment`, `pos` are the wide columns, expanded on the host by
    the library the first time one of them is read.  All arrays are zero-copy views of library-owned memory."""

    def __init__(self, ctx=None, hh=None, cols=None):
        self._cols = cols            # explicit wide columns (select() / empty lists)
        self._owner = None
        self._n = 0 if cols is None else int(cols[3].size)
        self._wide = None
        self._compact = None
        if hh is not None:
            self._owner = _HitsHandle(ctx, hh)
            self._n = int(ctx.L.pwmscan_hits_count(hh))

    def __len__(self):
        return self._n

    def first_invalid(self, alphabet="IUPAC"):
        """fromBS verdict of the bytes a host scan uploaded for this list (pwmscan_hits_first_invalid): offset of the first
        byte outside the alphabet, -1 if none (and for scans of resident sequences)."""
        if self._owner is None:
            return -1
        pos = C.c_int64(-1)
        ctx = self._owner.ctx
        ctx.check(ctx.L.pwmscan_hits_first_invalid(self._owner.h, C.c_int(_alphabet_code(alphabet)), C.byref(pos)))
        return int(pos.value)

    def _get_compact(self):
        if self._compact is None:
            o, n = self._owner, self._n
            p32, psc, pro, psg, pms = (C.POINTER(C.c_uint32)(), C.POINTER(C.c_double)(), C.POINTER(C.c_int64)(),
                                       C.POINTER(C.c_int32)(), C.POINTER(C.c_uint16)())
            o.L.pwmscan_hits_compact(o.h, C.byref(p32), C.byref(psc), C.byref(pro), C.byref(psg), C.byref(pms))
            self._compact = (_view(p32, n, o), _view(psc, n, o),
                             _view(pro, 2 * o.M + 1, o) if pro else None,
                             _view(psg, n, o) if psg else None, _view(pms, n, o) if pms else None)
        return self._compact

    def _get_wide(self):
        if self._cols is not None:
            return self._cols
        if self._wide is None:
            if self._n == 0:
                self._wide = (np.zeros(0, np.int32), np.zeros(0, np.uint8), np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(0))
            else:
                o, n = self._owner, self._n
                pm, ps, pg, pp, pc = (C.POINTER(C.c_int32)(), C.POINTER(C.c_uint8)(), C.POINTER(C.c_int32)(),
                                      C.POINTER(C.c_int64)(), C.POINTER(C.c_double)())
                o.ctx.check(o.L.pwmscan_hits_columns(o.h, C.byref(pm), C.byref(ps), C.byref(pg), C.byref(pp), C.byref(pc)))
                self._wide = (_view(pm, n, o), _view(ps, n, o), _view(pg, n, o), _view(pp, n, o), _view(pc, n, o))
        return self._wide

    motif = property(lambda self: self._get_wide()[0])
    strand = property(lambda self: self._get_wide()[1])
    segment = property(lambda self: self._get_wide()[2])
    pos = property(lambda self: self._get_wide()[3])

    @property
    def score(self):
        if self._cols is not None or self._n == 0:
            return self._get_wide()[4]
        return self._get_compact()[1]

    @property
    def pos32(self):
        return self._get_compact()[0] if self._n and self._cols is None else self.pos.astype(np.uint32)

    @property
    def run_off(self):
        """Single-segment scans: int64[2M+1] run table of (motif, strand); None otherwise."""
        return self._get_compact()[2] if self._n and self._cols is None else None

    def affinity(self, cdf_off, cdf_x, cdf_y, want_pvalue=False):
        """scanMotif's per-hit BED score (and p-value) on the device: pwmscan_hits_affinity.  cdf_off[M+1]
        indexes the concatenated (score, cumulative probability) points of every motif's truncated CDF."""
        n = len(self)
        aff = np.empty(n, dtype=np.int64)
        pv = np.empty(n) if want_pvalue else None
        owner = self._owner
        if n:
            if owner is None or not owner.h:
                raise ValueError("Hits.affinity needs the library-owned hit list (not a select()/copy)")
            off, cx, cy = _arr(cdf_off, np.int64), _arr(cdf_x, np.float64), _arr(cdf_y, np.float64)
            ctx = owner.ctx
            ctx.check(ctx.L.pwmscan_hits_affinity(ctx.h, owner.h, C.c_int32(off.size - 1), _ptr(off), _ptr(cx), _ptr(cy),
                                                  _ptr(pv) if want_pvalue else None, _ptr(aff)))
        return (aff, pv) if want_pvalue else aff

    def select(self, motif=None, strand=None, segment=None):
        k = np.ones(self._n, dtype=bool)
        if motif is not None:
            k &= self.motif == motif
        if strand is not None:
            k &= self.strand == strand
        if segment is not None:
            k &= self.segment == segment
        return Hits(cols=(self.motif[k], self.strand[k], self.segment[k], self.pos[k], self.score[k]))


_HITS_CB = C.CFUNCTYPE(None, _vp, C.c_int32, _vp)


class Plan:
    """Device-side analogue of [CutoffMotif] (pwmscan_plan): tables for fixed (motifs, cutoffs)."""

    def __init__(self, motifs, thres, strands=2, skip_ambiguous=True):
        self.ctx = motifs.ctx
        self.motifs = motifs
        th = _arr(thres, np.float64)
        if th.size != motifs.M:
            raise ValueError("one cutoff per motif expected")
        self.h = _vp()
        self.ctx.check(self.ctx.L.pwmscan_plan_create(self.ctx.h, motifs.h, _ptr(th), C.c_int(strands),
                                                      C.c_int(1 if skip_ambiguous else 0), C.byref(self.h)))
        self.ctx._retain()

    def info(self):
        out = np.zeros(8)
        self.ctx.L.pwmscan_plan_info(self.h, _ptr(out), C.c_int(8))
        mode = ("lds", "mma", "kmer")[int(out[7])]
        d = {"mode": mode, "blocks": int(out[0]), "exact_combos": int(out[2]), "candidates_per_position": out[3],
             "table_bytes": out[4], "genome_passes": out[5]}
        if mode == "kmer":       # k-mer mask tables gathered from L2
            d["table_sectors_per_position"] = out[1]
            d["first_stage_survivors_per_position"] = out[6]
        elif mode == "lds":      # 4-mer deficit tables in shared memory
            d["smem_wavefronts_per_position"] = out[1]
            d["p_second_stage_sum"] = out[6]
        return d

    def scan_host(self, data, seg_off=None, uppercase=False):
        """One-shot scan of host bytes (pwmscan_scan_host): upload, packing and scan are pipelined."""
        if isinstance(data, (bytes, bytearray, memoryview)):
            data = np.frombuffer(bytes(data), dtype=np.uint8)
        data = _arr(data, np.uint8)
        so = _arr(seg_off if seg_off is not None else [0, data.size], np.int64)
        hh = _vp()
        self.ctx.check(self.ctx.L.pwmscan_scan_host(self.ctx.h, self.h, _ptr(data), _ptr(so), C.c_int32(so.size - 1),
                                                   C.c_int(SEQ_UPPERCASE if uppercase else 0), C.byref(hh)))
        return self._hits(hh)

    def scan(self, seq):
        hh = _vp()
        self.ctx.check(self.ctx.L.pwmscan_scan(self.ctx.h, seq.h, self.h, C.byref(hh)))
        return self._hits(hh)

    def _hits(self, hh):
        h = Hits(self.ctx, hh)
        h._owner.M = self.motifs.M
        return h

    def _many(self, call, n, consume):
        out = [None] * n
        err = []
        c = consume or (lambda h: h)

        def cb(_user, i, hh):
            try:
                out[i] = c(self._hits(_vp(hh)))
            except BaseException as e:          # never let an exception cross the C frame
                err.append(e)
        fn = _HITS_CB(cb)
        self.ctx.check(call(fn))
        if err:
            raise err[0]
        return out

    def scan_many(self, seqs, consume=None):
        """pwmscan_scan_many: the items (DeviceSeq) are scanned with several of them in flight on the device.  Returns
        [Hits per item], or [consume(hits)] — the consumer runs in item order as soon as an item's hits have landed, and
        a hit list that is dropped goes straight back to the pinned pool (a streaming consumer, like scan_motif's writer)."""
        arr = (_vp * len(seqs))(*[s.h for s in seqs])
        return self._many(lambda fn: self.ctx.L.pwmscan_scan_many(self.ctx.h, self.h, C.c_int32(len(seqs)), arr, fn, None), len(seqs), consume)

    def scan_host_many(self, arrays, consume=None, uppercase=False):
        """pwmscan_scan_host_many: host byte arrays (pinned for full PCIe speed), one single-segment scan each."""
        arrays = [_arr(a, np.uint8) for a in arrays]
        ptrs = (_vp * len(arrays))(*[a.ctypes.data for a in arrays])
        lens = np.array([a.size for a in arrays], dtype=np.int64)
        return self._many(lambda fn: self.ctx.L.pwmscan_scan_host_many(self.ctx.h, self.h, C.c_int32(len(arrays)), ptrs, _ptr(lens),
                                                                        C.c_int(SEQ_UPPERCASE if uppercase else 0), fn, None),
                          len(arrays), consume)

    def scan_packed_many(self, items, consume=None):
        """pwmscan_scan_packed_many: items = [(PackedSeq, start, end)] — slices of the packed genome store in host memory
        (pinned for full PCIe speed), 0.375 bytes per base over PCIe instead of 1."""
        views = [p.slice(st, en) for p, st, en in items]
        p2 = (_vp * len(views))(*[v[0].ctypes.data for v in views])
        pm = (_vp * len(views))(*[v[1].ctypes.data for v in views])
        lens = np.array([v[2] for v in views], dtype=np.int64)
        return self._many(lambda fn: self.ctx.L.pwmscan_scan_packed_many(self.ctx.h, self.h, C.c_int32(len(views)), p2, pm, _ptr(lens), fn, None),
                          len(views), consume)

    def free(self):
        if self.h:
            self.ctx.L.pwmscan_plan_free(self.h)
            self.h = _vp()
            self.ctx._release()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def scores(seq, motifs, m, strand=0):
    n = int(motifs.ncol[m])
    nwin = max(0, seq.length - n + 1)
    out = np.empty(nwin)
    seq.ctx.check(seq.ctx.L.pwmscan_scores(seq.ctx.h, seq.h, motifs.h, C.c_int32(m), C.c_int(strand), _ptr(out), C.c_int64(nwin)))
    return out


def max_match_sc(seq, motifs):
    out = np.empty(motifs.M)
    seq.ctx.check(seq.ctx.L.pwmscan_max_match_sc(seq.ctx.h, seq.h, motifs.h, _ptr(out)))
    return out


def exported_symbols():
    """Names declared in include/pwmscan.h (used by the CPU test that checks the .so exports them all)."""
    import re
    hdr = os.path.join(_HERE, "..", "include", "pwmscan.h")
    txt = open(hdr).read()
    return sorted(set(re.findall(r"\b(pwmscan_[a-z0-9_]+)\s*\(", txt)))
