"""Bio.Seq.IO mirror (bioinformatics-toolkit/src/Bio/Seq/IO.hs): the flat indexed genome file.

File layout (SeqIO.hs:33-34,43-48,96-112): magic line, one header line of `name offset length`
triples, then the raw concatenated sequence bytes exactly as found in the FASTA (case preserved).
Host-side I/O; the bytes go to the device through `_lib.DeviceSeq` (which upper-cases and validates
like fromBS when asked to).
"""
import os

import numpy as np

from .seq import DNA, fromBS

MAGIC = b"<HASKELLBIOINFORMATICS_7d2c5gxhg934>"


class Genome:
    """data Genome = Genome !Handle !IndexTable !Int (SeqIO.hs:29)."""

    def __init__(self, path):
        self.path = path
        with open(path, "rb") as f:
            sig = f.readline().rstrip(b"\n")
            if sig != MAGIC:
                raise ValueError("Bio.Seq.Query.openGenome: Incorrect format")
            header = f.readline().rstrip(b"\n")
        w = header.split()
        if len(w) % 3:
            raise ValueError("error")
        self.index = {}
        self.order = []
        for i in range(0, len(w), 3):
            self.index[w[i]] = (int(w[i + 1]), int(w[i + 2]))
            self.order.append(w[i])
        self.header_size = len(sig) + len(header) + 2
        self._mm = np.memmap(path, dtype=np.uint8, mode="r")
        self.data = self._mm.view(np.ndarray)          # plain ndarray view: slicing a np.memmap object is ~20x slower

    def close(self):
        self.data = None
        self._mm = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def openGenome(path):
    return Genome(path)


def closeGenome(g):
    g.close()


def withGenome(path, fn):
    with Genome(path) as g:
        return fn(g)


def raw_bytes(g, query):
    """(ok, bytes-or-message): the un-validated bytes of a query, or the reference's Left message."""
    chr_, start, end = query
    if isinstance(chr_, str):
        chr_ = chr_.encode()
    if chr_ not in g.index:
        return False, "Bio.Seq.getSeq: Cannot find " + repr(chr_.decode("latin1"))
    cstart, csize = g.index[chr_]
    if end > csize:
        return False, "Bio.Seq.getSeq: out of index: %d>%d" % (end, csize)
    o = g.header_size + cstart + start
    return True, g.data[o:o + max(0, end - start)]


def getSeq(g, query, alphabet="IUPAC"):
    """(ok, DNA | message) — SeqIO.hs:65-75; zero-based half-open (chr, start, end)."""
    ok, r = raw_bytes(g, query)
    if not ok:
        return False, r
    return fromBS(r.tobytes(), alphabet)


def getChrom(g, chr_):
    if isinstance(chr_, str):
        chr_ = chr_.encode()
    if chr_ not in g.index:
        return False, "Unknown chromosome"
    return getSeq(g, (chr_, 0, g.index[chr_][1]))


def getChrSizes(g):
    return [(k, g.index[k][1]) for k in g.order]


def fastaReader(path):
    """Yields (header without '>', [sequence lines]) — Data/Fasta.hs:37-51."""
    name, acc = None, []
    with open(path, "rb") as f:
        for l in f:
            l = l.rstrip(b"\n").rstrip(b"\r")
            if not l:
                continue
            if l[:1] == b">":
                if name is not None:
                    yield name, acc
                name, acc = l[1:], []
            elif name is not None:
                acc.append(l)
    if name is not None:
        yield name, acc


def mkIndex(fasta_files, out_path):
    """SeqIO.hs:92-112."""
    chrs, dnas = [], []
    for fl in fasta_files:
        for name, lines in fastaReader(fl):
            dna = b"".join(lines)
            chrs.append((name.split()[0], len(dna)))
            dnas.append(dna)
    hdr, off = [], 0
    for nm, ln in chrs:
        hdr += [nm, str(off).encode(), str(ln).encode()]
        off += ln
    with open(out_path, "wb") as f:
        f.write(MAGIC + b"\n" + b" ".join(hdr) + b"\n")
        for d in dnas:
            f.write(d)


def write_genome(chroms, out_path):
    """Index file straight from [(name, uint8 array)] (synthetic genomes in tests / bench)."""
    hdr, off = [], 0
    for nm, arr in chroms:
        hdr += [nm if isinstance(nm, bytes) else nm.encode(), str(off).encode(), str(len(arr)).encode()]
        off += len(arr)
    with open(out_path, "wb") as f:
        f.write(MAGIC + b"\n" + b" ".join(hdr) + b"\n")
        for _, arr in chroms:
            f.write(np.ascontiguousarray(arr, dtype=np.uint8).tobytes())


# ---- the packed genome store next to the index file -------------------------------------------------------------------
# `<genome>.packed`: the 2-bit + invalid-base planes of every chromosome (csrc/hostpack.cu, 0.375 bytes per base), made
# ONCE at index time (upper-cased like fromBS, Seq.hs:86-95) so that whole-genome scans upload those instead of the bytes.
# Layout: magic line, one JSON line [{name, length, first_non_acgt, first_non_iupac, seq2_off, mask_off}] (offsets in bytes
# from the start of the data section, 64-byte aligned), then the uint32 planes.
PACKED_MAGIC = b"<PWMSCAN_PACKED_GENOME_V1>"


def mkPackedIndex(genome, out_path=None, nthreads=0):
    """Writes `<genome.path>.packed` (or out_path) from an open Genome.  Returns the path."""
    import json
    from . import _lib
    out_path = out_path or genome.path + ".packed"
    meta, planes, off = [], [], 0
    for name in genome.order:
        cstart, csize = genome.index[name]
        o = genome.header_size + cstart
        p = _lib.PackedSeq(genome.data[o:o + csize], uppercase=True, nthreads=nthreads)
        rec = {"name": name.decode("latin1"), "length": csize, "first_non_acgt": p.first_non_acgt, "first_non_iupac": p.first_non_iupac}
        for key, arr in (("seq2_off", p.seq2), ("mask_off", p.mask)):
            rec[key] = off
            planes.append(arr)
            off += (arr.nbytes + 63) // 64 * 64
        meta.append(rec)
    with open(out_path, "wb") as f:
        f.write(PACKED_MAGIC + b"\n" + json.dumps(meta).encode() + b"\n")
        for arr in planes:
            f.write(arr.tobytes())
            f.write(b"\0" * ((-arr.nbytes) % 64))
    return out_path


class PackedGenome:
    """The `.packed` sidecar, memory-mapped: `planes(chr)` -> _lib.PackedPlanes of a chromosome (views into the file)."""

    def __init__(self, path):
        import json
        from . import _lib
        with open(path, "rb") as f:
            if f.readline().rstrip(b"\n") != PACKED_MAGIC:
                raise ValueError("not a packed genome store: " + path)
            meta = json.loads(f.readline())
            base = f.tell()
        self._mm = np.memmap(path, dtype=np.uint8, mode="r")
        data = self._mm.view(np.ndarray)
        self.chroms = {}
        for rec in meta:
            units = (rec["length"] + 31) // 32
            s2 = data[base + rec["seq2_off"]: base + rec["seq2_off"] + 8 * units].view(np.uint32)
            mk = data[base + rec["mask_off"]: base + rec["mask_off"] + 4 * units].view(np.uint32)
            self.chroms[rec["name"].encode("latin1")] = _lib.PackedPlanes(s2, mk, rec["length"], rec["first_non_acgt"], rec["first_non_iupac"])

    def planes(self, chr_):
        return self.chroms[chr_ if isinstance(chr_, bytes) else chr_.encode()]

    def close(self):
        self.chroms = {}
        self._mm = None
