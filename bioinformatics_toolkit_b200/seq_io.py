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
