"""Bio.Data.Bed.Utils motif part (bioinformatics-toolkit/src/Bio/Data/Bed/Utils.hs:76-124).

CutoffMotif (78-86), mkCutoffMotif (88-98), scanMotif (101-124).  scanMotif batches the BED
regions: the region bytes are gathered from the genome file, laid end to end (one segment per
region), packed and scanned on the device in one launch per batch for all motifs and both strands;
hits come back in the reference's order (region, motif, forward hits ascending, reverse hits
ascending).  The per-hit BED score (toAffinity of 1 - cdf score) is computed on the device too
(pwmscan_hits_affinity); `_cdf_vec` / `_affinity_vec` are the host restatement the tests compare it with.
"""
import math
import os
import sys
import time

import numpy as np

from . import _lib
from .motif import CDF, rcPWM, size
from .seq import ALPHABET, _UPPER
from .seq_io import raw_bytes


class CutoffMotif:
    __slots__ = ("_motif_name", "_motif_pwm", "_motif_sigma", "_motif_pwm_rc", "_motif_sigma_rc", "_background", "_cutoff", "_cdf")

    def __init__(self, name, pwm, sigma, pwm_rc, sigma_rc, bg, cutoff, cdf):
        self._motif_name, self._motif_pwm, self._motif_sigma = name, pwm, sigma
        self._motif_pwm_rc, self._motif_sigma_rc, self._background, self._cutoff, self._cdf = pwm_rc, sigma_rc, bg, cutoff, cdf


def mkCutoffMotifs(bg, p, motifs, ctx=None):
    """[mkCutoffMotif bg p m | m <- motifs], with all score CDFs computed in batched launches."""
    ctx = ctx or _lib.default_context()
    mo = _lib.DeviceMotifs(ctx, [m._pwm._mat for m in motifs], bg.freqs)
    cut, cdfs = mo.cutoff_cdfs(p)
    out = []
    for i, m in enumerate(motifs):
        out.append(CutoffMotif(m._name, m._pwm, mo.get_sigma(i, 0), rcPWM(m._pwm), mo.get_sigma(i, 1), bg, float(cut[i]), CDF(*cdfs[i])))
    return out


def mkCutoffMotif(bg, p, motif):
    return mkCutoffMotifs(bg, p, [motif])[0]


def _cdf_vec(c, x):
    """cdf (Motif.hs:236-249) for an array of scores; elementwise IEEE arithmetic, no transcendentals."""
    xs, ps = c.xs, c.ps
    n = xs.size
    i = np.searchsorted(xs, x, side="left")
    out = np.empty(x.size)
    hi = i >= n
    lo = (i == 0) & ~hi
    mid = ~hi & ~lo
    out[hi] = ps[-1] if n else 0.0
    out[lo] = np.where(x[lo] == (xs[0] if n else 0.0), ps[0] if n else 0.0, 0.0)
    im = i[mid]
    a, a1, b, b1 = xs[im - 1], ps[im - 1], xs[im], ps[im]
    al = (b - x[mid]) / (b - a)
    be = (x[mid] - a) / (b - a)
    out[mid] = al * a1 + be * b1
    return out


def toAffinity(x1):
    """round (1000 / (1 + exp (-(-log10 (max 1e-20 x') - 5)))), half-even, libm (Utils.hs:120-123)."""
    x = -(math.log(max(1e-20, x1)) / math.log(10.0))
    sc = 1 / (1 + math.exp(-(x - 5)))
    return int(round(sc * 1000))       # Python round() on a float is round-half-even, like Haskell's


def _affinity_vec(pv):
    x = -(np.log(np.maximum(1e-20, pv)) / math.log(10.0))
    v = 1 / (1 + np.exp(-(x - 5))) * 1000
    out = np.rint(v).astype(np.int64)
    # numpy's SIMD log/exp may differ from libm in the last ulp: settle anything near a rounding tie with libm
    near = np.nonzero(np.abs(np.abs(v - np.floor(v)) - 0.5) < 1e-6)[0]
    for k in near:
        out[k] = toAffinity(float(pv[k]))
    return out


def scanMotif(genome, motifs, beds, out=None, batch_bases=192 << 20, warn=sys.stderr, ctx=None):
    """beds: iterable of (chrom, start, end).  Yields BED6 tuples
    (chrom, start, end, name, score, strand) — or, if `out` is a binary file object, writes the
    reference's `toLine` text (Data/Bed/Types.hs:139-149) to it and returns the hit count."""
    ctx = ctx or _lib.default_context()
    if not motifs:
        return 0 if out is not None else []
    bg = motifs[0]._background
    trace = os.environ.get("PWMSCAN_APP_TIMING") is not None
    t_start = time.perf_counter()
    mo = _lib.DeviceMotifs(ctx, [m._motif_pwm._mat for m in motifs], bg.freqs)
    for i, m in enumerate(motifs):                                 # findTFBSWith takes the stored sigmas
        mo.set_sigma(i, 0, m._motif_sigma)
        mo.set_sigma(i, 1, m._motif_sigma_rc)
    plan = _lib.Plan(mo, [m._cutoff for m in motifs], strands=2, skip_ambiguous=True)
    if trace:
        print("[scanMotif] motifs + plan %.1f ms" % (1e3 * (time.perf_counter() - t_start)), file=sys.stderr)
    nlen = np.array([size(m._motif_pwm) for m in motifs], dtype=np.int64)
    names = [m._motif_name for m in motifs]
    valid = np.zeros(256, dtype=bool)
    valid[np.frombuffer(ALPHABET["IUPAC"], dtype=np.uint8)] = True
    results = [] if out is None else None
    total = 0

    import ctypes as C
    L = ctx.L
    mname_off = np.concatenate([[0], np.cumsum([len(x) for x in names])]).astype(np.int64)
    mname_buf = b"".join(names)
    mlen32 = nlen.astype(np.int32)

    # every motif's truncated CDF, concatenated, for the device-side per-hit p-value (pwmscan_hits_affinity)
    cdf_off = np.concatenate([[0], np.cumsum([m._cdf.xs.size for m in motifs])]).astype(np.int64)
    cdf_x = np.concatenate([m._cdf.xs for m in motifs]) if motifs else np.zeros(0)
    cdf_y = np.concatenate([m._cdf.ps for m in motifs]) if motifs else np.zeros(0)

    def submit(regs, chunks):
        """One batch of regions: a single pipelined scan of the concatenated bytes (one segment per region),
        per-hit scores on the device, BED text in the library."""
        nonlocal total
        if not regs:
            return
        t_sub = time.perf_counter()
        lens = np.array([c.size for c in chunks], dtype=np.int64)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        data = np.concatenate(chunks)
        t_scan = time.perf_counter()
        h = plan.scan_host(data, seg_off=off, uppercase=True)
        n = len(h)
        if trace:
            print("[scanMotif] concatenate %.1f ms; scan_host %d regions, %d bases, %d hits: %.1f ms"
                  % (1e3 * (t_scan - t_sub), len(regs), data.size, n, 1e3 * (time.perf_counter() - t_scan)), file=sys.stderr)
        # fromBS per region (Seq.hs:86-95): a region with a byte outside the IUPAC alphabet is a Left (warning,
        # skipped).  The pack kernel has already looked at every byte; only if it saw one is the batch examined
        # region by region on the host.
        keep = None
        if data.size and ctx.scan_host_first_invalid("IUPAC") >= 0:
            bad = ~valid[_UPPER[data]]
            nz = lens > 0
            nbad = np.zeros(len(regs), dtype=np.int64)
            nbad[nz] = np.add.reduceat(bad.astype(np.int64), off[:-1][nz])
            bad_regions = nbad > 0
            if warn is not None:
                for r in (regs[i] for i in np.flatnonzero(bad_regions)):
                    print("Warning: no sequence for region: " + repr((r[0].decode("latin1"), r[1], r[2])), file=warn)
            keep = ~bad_regions[h.segment]
        if n == 0:
            return
        # per-hit p-value (1 - cdf score on the motif's truncated CDF) and affinity, on the device
        t_pv = time.perf_counter()
        score = h.affinity(cdf_off, cdf_x, cdf_y)
        seg, mot, strand, pos = h.segment, h.motif, h.strand, h.pos
        if keep is not None:
            seg, mot, strand, pos, score = seg[keep], mot[keep], strand[keep], pos[keep], score[keep]
            n = int(seg.size)
            if n == 0:
                return
        total += n
        if trace:
            print("[scanMotif] p-values + affinities: %.1f ms" % (1e3 * (time.perf_counter() - t_pv)), file=sys.stderr)
        t_fmt = time.perf_counter()
        starts = np.array([r[1] for r in regs], dtype=np.int64)
        if out is None:
            st = starts[seg] + pos
            en = st + nlen[mot]
            for k in range(n):
                results.append((regs[seg[k]][0], int(st[k]), int(en[k]), names[mot[k]], int(score[k]), strand[k] == 0))
            return
        chroms = sorted(set(r[0] for r in regs))
        cidx = {c: i for i, c in enumerate(chroms)}
        rchrom = np.array([cidx[r[0]] for r in regs], dtype=np.int32)
        coff = np.concatenate([[0], np.cumsum([len(c) for c in chroms])]).astype(np.int64)
        cbuf = b"".join(chroms)
        text = C.c_char_p()
        tlen = C.c_int64(0)
        p = _lib._ptr
        rc = L.pwmscan_format_bed6(C.c_int64(n), p(np.ascontiguousarray(seg)), p(np.ascontiguousarray(mot)),
                                   p(np.ascontiguousarray(strand)), p(np.ascontiguousarray(pos)), p(score), p(rchrom), p(starts),
                                   C.c_char_p(cbuf), p(coff), C.c_char_p(mname_buf), p(mname_off), p(mlen32),
                                   C.byref(text), C.byref(tlen))
        if rc != 0:
            raise _lib.PwmScanError(rc, "pwmscan_format_bed6 failed")
        out.write(memoryview((C.c_char * tlen.value).from_address(C.cast(text, C.c_void_p).value)))   # no intermediate bytes object
        L.pwmscan_free(text)
        if trace:
            print("[scanMotif] BED text: %.1f ms" % (1e3 * (time.perf_counter() - t_fmt)), file=sys.stderr)

    index, hsize, gdata = genome.index, genome.header_size, genome.data
    regs, chunks, nb = [], [], 0
    t_loop = time.perf_counter()
    for (chr_, start, end) in beds:
        if isinstance(chr_, str):
            chr_ = chr_.encode()
        ent = index.get(chr_)
        if ent is None or end > ent[1]:                            # getSeq: Left (SeqIO.hs:65-75)
            if warn is not None:
                print("Warning: no sequence for region: " + repr((chr_.decode("latin1"), start, end)), file=warn)
            continue
        o = hsize + ent[0] + start
        regs.append((chr_, start, end))
        chunks.append(gdata[o:o + max(0, end - start)])
        nb += max(0, end - start)
        if nb >= batch_bases or len(regs) >= 65000:
            submit(regs, chunks)
            regs, chunks, nb = [], [], 0
    if trace:
        print("[scanMotif] BED loop + region views: %.1f ms (includes earlier batches)" % (1e3 * (time.perf_counter() - t_loop)), file=sys.stderr)
    submit(regs, chunks)
    # release the device tables now (hundreds of MiB for 800 motifs) rather than whenever the collector gets to them
    t_free = time.perf_counter()
    plan.free()
    mo.free()
    if trace:
        print("[scanMotif] free %.1f ms, total %.1f ms" % (1e3 * (time.perf_counter() - t_free), 1e3 * (time.perf_counter() - t_start)), file=sys.stderr)
    return total if out is not None else results


def scanGenome(genome, motifs, out=None, chroms=None, chunk_len=64_000_000, warn=sys.stderr, ctx=None, packed=None):
    """scanMotif over WHOLE chromosomes straight from the genome store — the result of
    `scanMotif genome motifs [(chr, 0, size) | (chr, size) <- getChrSizes genome]` (Utils.hs:101-124 with getChrom,
    Seq/IO.hs:77-84), produced chromosome by chromosome: the bytes of a chromosome are read from the index file in
    chunks of `chunk_len` window starts (+ a halo of the longest motif - 1), go through pinned staging buffers into
    pwmscan_scan_host_many (upload, 2-bit packing, first stage, exact replay and hit copy of several chunks in flight), the
    per-hit BED score comes from the device (pwmscan_hits_affinity) and the chunks' (motif, strand) runs are stitched
    into the reference's order (motif, strand, ascending position over the whole chromosome).  A chromosome with a byte
    outside the IUPAC alphabet is a Left in the reference (fromBS): warned about and skipped.
    With `packed` (a seq_io.PackedGenome: the `.packed` sidecar of the index, `mkindex --packed`) the chunks are slices of
    the chromosome's 2-bit + mask planes instead (0.375 bytes per base through the staging buffers and over PCIe,
    pwmscan_scan_packed_many; chunk_len is rounded down to a multiple of 32) — same output.
    Yields / writes BED6 like scanMotif; returns the hit count when `out` is given."""
    import ctypes as C
    import torch
    from . import shard
    ctx = ctx or _lib.default_context()
    if not motifs:
        return 0 if out is not None else []
    bg = motifs[0]._background
    mo = _lib.DeviceMotifs(ctx, [m._motif_pwm._mat for m in motifs], bg.freqs)
    for i, m in enumerate(motifs):
        mo.set_sigma(i, 0, m._motif_sigma)
        mo.set_sigma(i, 1, m._motif_sigma_rc)
    plan = _lib.Plan(mo, [m._cutoff for m in motifs], strands=2, skip_ambiguous=True)
    M = len(motifs)
    nlen = np.array([size(m._motif_pwm) for m in motifs], dtype=np.int64)
    nmax = int(nlen.max())
    names = [m._motif_name for m in motifs]
    mname_off = np.concatenate([[0], np.cumsum([len(x) for x in names])]).astype(np.int64)
    mname_buf = b"".join(names)
    mlen32 = nlen.astype(np.int32)
    cdf_off = np.concatenate([[0], np.cumsum([m._cdf.xs.size for m in motifs])]).astype(np.int64)
    cdf_x = np.concatenate([m._cdf.xs for m in motifs])
    cdf_y = np.concatenate([m._cdf.ps for m in motifs])
    L = ctx.L
    p = _lib._ptr
    results = [] if out is None else None
    total = 0
    staging = []                                    # pinned host buffers, reused from chromosome to chromosome
    for chr_, csize in (getChrSizesOf(genome) if chroms is None else [(c if isinstance(c, bytes) else c.encode(), genome.index[c if isinstance(c, bytes) else c.encode()][1]) for c in chroms]):
        cstart = genome.index[chr_][0]
        if packed is not None:
            pl = packed.planes(chr_)
            if pl.first_non_iupac >= 0:                                 # fromBS's verdict, taken when the store was packed
                if warn is not None:
                    print("Warning: no sequence for region: " + repr((chr_.decode("latin1"), 0, csize)), file=warn)
                continue
            items = shard.make_items([csize], M, max(32, chunk_len // 32 * 32))
            pitems = []
            for k, it in enumerate(items):
                lo, hi = shard.chunk_bytes_range(it, csize, nmax)
                s2, mk, n = pl.slice(lo, hi)
                units = (n + 31) // 32
                while len(staging) <= k:
                    staging.append(None)
                if staging[k] is None or not isinstance(staging[k], tuple) or staging[k][0].numel() < 2 * units:
                    staging[k] = (torch.empty(max(2 * units, 2), dtype=torch.int32).pin_memory(), torch.empty(max(units, 1), dtype=torch.int32).pin_memory())
                a2 = staging[k][0].numpy().view(np.uint32)[:2 * units]
                am = staging[k][1].numpy().view(np.uint32)[:units]
                np.copyto(a2, s2[:2 * units])                           # sidecar (page cache) -> pinned memory
                np.copyto(am, mk[:units])
                pitems.append((_lib.PackedPlanes(a2, am, n), 0, None))
            hits = plan.scan_packed_many(pitems)
        items = items if packed is not None else shard.make_items([csize], M, chunk_len)
        arrays = []
        for k, it in enumerate(items if packed is None else []):
            lo, hi = shard.chunk_bytes_range(it, csize, nmax)
            if k >= len(staging) or staging[k].numel() < hi - lo:
                buf = torch.empty(max(hi - lo, 1), dtype=torch.uint8).pin_memory()
                if k < len(staging):
                    staging[k] = buf
                else:
                    staging.append(buf)
            a = staging[k].numpy()[:hi - lo]
            o = genome.header_size + cstart + lo
            np.copyto(a, genome.data[o:o + (hi - lo)])              # index file (page cache) -> pinned memory
            arrays.append(a)
        if packed is None:
            hits = plan.scan_host_many(arrays, uppercase=True)
        if packed is None and any(h.first_invalid("IUPAC") >= 0 for h in hits):        # fromBS: Left "Bio.Seq.fromBS: unknown character"
            if warn is not None:
                print("Warning: no sequence for region: " + repr((chr_.decode("latin1"), 0, csize)), file=warn)
            continue
        # chunk c contributes the windows that START in it; its runs of (motif, strand) go behind those of chunk c - 1
        parts = []
        for it, h in zip(items, hits):
            n = len(h)
            if n == 0:
                continue
            own = it[2] - it[1]
            pos = h.pos32.astype(np.int64)
            aff = h.affinity(cdf_off, cdf_x, cdf_y)
            ro = np.asarray(h.run_off)
            combo = np.repeat(np.arange(2 * M, dtype=np.int64), np.diff(ro))
            keep = pos < own
            parts.append((combo[keep], pos[keep] + it[1], aff[keep]))
        if not parts:
            continue
        combo = np.concatenate([q[0] for q in parts])
        pos = np.concatenate([q[1] for q in parts])
        aff = np.concatenate([q[2] for q in parts])
        order = np.argsort(combo, kind="stable")                    # chunks are in position order already
        combo, pos, aff = combo[order], pos[order], aff[order]
        mot = (combo >> 1).astype(np.int32)
        strand = (combo & 1).astype(np.uint8)
        n = int(pos.size)
        total += n
        if out is None:
            en = pos + nlen[mot]
            for k in range(n):
                results.append((chr_, int(pos[k]), int(en[k]), names[mot[k]], int(aff[k]), strand[k] == 0))
            continue
        text = C.c_char_p()
        tlen = C.c_int64(0)
        seg = np.zeros(n, dtype=np.int32)
        rchrom = np.zeros(1, dtype=np.int32)
        starts = np.zeros(1, dtype=np.int64)
        coff = np.array([0, len(chr_)], dtype=np.int64)
        rc = L.pwmscan_format_bed6(C.c_int64(n), p(seg), p(mot), p(strand), p(np.ascontiguousarray(pos)), p(np.ascontiguousarray(aff)), p(rchrom), p(starts),
                                   C.c_char_p(chr_), p(coff), C.c_char_p(mname_buf), p(mname_off), p(mlen32), C.byref(text), C.byref(tlen))
        if rc != 0:
            raise _lib.PwmScanError(rc, "pwmscan_format_bed6 failed")
        out.write(memoryview((C.c_char * tlen.value).from_address(C.cast(text, C.c_void_p).value)))
        L.pwmscan_free(text)
    plan.free()
    mo.free()
    return total if out is not None else results


def getChrSizesOf(genome):
    return [(k, genome.index[k][1]) for k in genome.order]


def streamBed3(path):
    """streamBed ... :: BED3 (Data/Bed.hs:307-309, Types.hs:175-177): first three tab-separated fields."""
    with open(path, "rb") as f:
        for l in f:
            l = l.rstrip(b"\n")
            if not l:
                continue
            w = l.split(b"\t")
            if len(w) < 3:
                raise ValueError("Read BED fail: Incorrect number of fields")
            yield w[0], int(w[1]), int(w[2])
