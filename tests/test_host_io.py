"""Host logic above the C ABI that needs no device: the genome index (Bio.Seq.IO, SeqIO.hs:29-112; fastaReader,
Fasta.hs:37-51), Bio.Seq (Seq.hs:41-119), motif files (Motif.hs:329-386), toIUPAC (107-124), rcPWM (77-78) and the
reference's own motif fixture read through these parsers."""
import os

import numpy as np
import pytest

from bioinformatics_toolkit_b200 import motif as Mo
from bioinformatics_toolkit_b200 import seq as Sq
from bioinformatics_toolkit_b200 import synth
from bioinformatics_toolkit_b200.seq_io import (closeGenome, fastaReader, getChrom, getChrSizes, getSeq, mkIndex, openGenome,
                                               write_genome)


def test_mkindex_from_fasta_and_getseq(tmp_path):
    fa1 = os.path.join(tmp_path, "a.fa")
    fa2 = os.path.join(tmp_path, "b.fa")
    c1 = bytes(synth.synth_dna(1234, 1, lower_frac=0.3))
    c2 = b"ACGTNNNNRYKMacgtnn"
    c3 = bytes(synth.synth_dna(77, 2))
    with open(fa1, "wb") as f:                                   # wrapped lines, a description after the name, CRLF, blank lines
        f.write(b">chr1 some description\n" + b"\n".join(c1[i:i + 60] for i in range(0, len(c1), 60)) + b"\n\n>chr2\r\n" + c2 + b"\r\n")
    with open(fa2, "wb") as f:
        f.write(b">chrZ\n" + c3 + b"\n")
    assert [n for n, _ in fastaReader(fa1)] == [b"chr1 some description", b"chr2"]
    idx = os.path.join(tmp_path, "g.idx")
    mkIndex([fa1, fa2], idx)
    g = openGenome(idx)
    assert getChrSizes(g) == [(b"chr1", len(c1)), (b"chr2", len(c2)), (b"chrZ", len(c3))]
    ok, dna = getSeq(g, (b"chr1", 10, 200))
    assert ok and Sq.toBS(dna) == c1[10:200].upper() and Sq.length(dna) == 190
    ok, dna = getSeq(g, (b"chr2", 0, len(c2)))
    assert ok and Sq.toBS(dna) == c2.upper()                     # IUPAC letters are accepted, case is folded
    assert getSeq(g, (b"chr2", 0, len(c2)), alphabet="Basic")[0] is False    # ... but not as plain ACGT
    assert getSeq(g, (b"chr2", 5, len(c2) + 1))[0] is False      # beyond the end: Left
    assert getSeq(g, (b"nope", 0, 1))[0] is False
    ok, whole = getChrom(g, b"chrZ")
    assert ok and Sq.toBS(whole) == c3
    closeGenome(g)
    # write_genome produces the same file as mkIndex for the same sequences
    idx2 = os.path.join(tmp_path, "g2.idx")
    write_genome([(b"chr1", np.frombuffer(c1, np.uint8)), (b"chr2", np.frombuffer(c2, np.uint8)), (b"chrZ", np.frombuffer(c3, np.uint8))], idx2)
    assert open(idx, "rb").read() == open(idx2, "rb").read()


def test_seq_basics():
    ok, d = Sq.fromBS(b"acgtN")
    assert ok and Sq.toBS(d) == b"ACGTN"
    assert Sq.fromBS(b"ACGX")[0] is False
    assert Sq.fromBS(b"ACGN", alphabet="Basic")[0] is False
    assert Sq.toBS(Sq.rc(Sq.unsafeFromBS(b"AACGT"))) == b"ACGTT"
    assert Sq.toBS(Sq.slice(1, 3, Sq.unsafeFromBS(b"AACGT"))) == b"ACG"


def test_motif_files_roundtrip(tmp_path, ref_motifs, ref_motifs_meme):
    mats = [m for _, m in ref_motifs] + synth.synth_pwms(4, 3, nmin=5, nmax=9)
    mats.append(np.array([[1e-7, 0.9999997, 1e-7, 1e-7], [0.25, 0.25, 0.25, 0.25]]))     # exponent-form numbers
    motifs = [Mo.Motif(b"name %d" % i, Mo.PWM(17, m)) for i, m in enumerate(mats)]
    meme = os.path.join(tmp_path, "m.meme")
    Mo.writeMEME(meme, motifs)
    back = Mo.readMEME(meme)
    assert [m._name for m in back] == [m._name for m in motifs] and all(m._pwm._nSites == 17 for m in back)

    # toShortest prints the shortest digits that identify the double, but the reference READS with bytestring-lexing's
    # readExponential (motif.read_double), which is not correctly rounded for 17-digit mantissas: a written file reads back
    # exactly or one ulp away — and exactly what read_double makes of the printed text
    def back_ok(a, b):
        x, y = a._pwm._mat, b._pwm._mat
        exp = np.array([[Mo.read_double(Mo._shortest(v).encode()) for v in row] for row in y])
        return x.shape == y.shape and np.array_equal(x, exp) and np.all(np.abs(x - y) <= np.spacing(y))
    assert all(back_ok(a, b) for a, b in zip(back, motifs))
    assert np.array_equal(back[-1]._pwm._mat, motifs[-1]._pwm._mat)                                    # short decimals: exact
    fa = os.path.join(tmp_path, "m.fasta")
    Mo.writeFasta(fa, motifs)
    back = Mo.readFasta_(fa)
    assert all(back_ok(a, b) and a._name == b._name for a, b in zip(back, motifs))
    assert b"1e-7 0.9999997 1e-7 1e-7" in open(fa, "rb").read()                           # toShortest's exponent form
    # the reference's own MEME fixture parses to the matrices frozen from it
    assert len(ref_motifs_meme) > 0 and all(m.shape[1] == 4 for _, m in ref_motifs_meme)


def test_rc_iupac_and_subpwm(ref_motifs):
    for _, m in ref_motifs:
        p = Mo.PWM(None, m)
        r = Mo.rcPWM(p)
        assert np.array_equal(r._mat, m[::-1, ::-1]) and np.array_equal(Mo.rcPWM(r)._mat, m)     # Tests/Motif.hs: rc . rc = id
        # M.subMatrix takes inclusive corners, so the reference's `subPWM i l` returns l + 1 rows (Motif.hs:72-73): kept
        assert Mo.size(Mo.subPWM(1, 3, p)) == 4 and np.array_equal(Mo.subPWM(1, 3, p)._mat, m[1:5])
        assert len(Sq.toBS(Mo.toIUPAC(p))) == m.shape[0]
    p = Mo.PWM(None, np.array([[0.97, 0.01, 0.01, 0.01], [0.45, 0.45, 0.05, 0.05], [0.25, 0.25, 0.25, 0.25], [0.05, 0.05, 0.45, 0.45]]))
    assert Sq.toBS(Mo.toIUPAC(p)) == b"AMNK"
    assert Mo.toPWM([b"0.1 0.2 0.3 0.4", b"1 0 0 0"])._mat.tolist() == [[0.1, 0.2, 0.3, 0.4], [1.0, 0.0, 0.0, 0.0]]


def test_merge_helpers_hand_checked():
    """mergePWM / mergePWMWeighted / dilute / trim (Motif/Merge.hs:27-111) on cases worked out by hand, plus agreement
    of mergePWMWeighted with the oracle-side restatement on random inputs and shifts of both signs."""
    from bioinformatics_toolkit_b200 import merge as M
    from oracle import app_oracle
    a = Mo.PWM(None, np.array([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0]]))
    b = Mo.PWM(None, np.array([[0, 0, 0, 1.0], [0.5, 0.5, 0, 0]]))
    m = M.mergePWM(a, b, 2)                       # b starts at column 2 of a: overlap of one column, one new column
    assert m._mat.tolist() == [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0.5, 0.5], [0.5, 0.5, 0, 0]]
    m = M.mergePWM(a, b, -1)                      # negative shift: roles swapped, a starts at column 1 of b
    assert m._mat.tolist() == [[0, 0, 0, 1], [0.75, 0.25, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]]
    mw, ws = M.mergePWMWeighted((a, [3, 3, 3]), (b, [1, 1]), 2)
    assert ws == [3, 3, 4, 1] and mw._mat[2].tolist() == [0, 0, 0.75, 0.25]
    d = M.dilute((mw, ws))                        # columns supported by fewer sites are pulled towards 0.25
    assert np.allclose(d._mat[3], (1 * np.array([0.5, 0.5, 0, 0]) + 0.25 * 3) / 4) and np.array_equal(d._mat[2], mw._mat[2])
    flat = np.array([[0.25, 0.25, 0.25, 0.25], [0.9, 0.05, 0.03, 0.02], [0.3, 0.3, 0.2, 0.2], [0.26, 0.24, 0.25, 0.25]])
    t = M.trim(Mo.default_bkgd, 0.1, Mo.PWM(None, flat))
    assert t._mat.tolist() == flat[1:2].tolist()            # uninformative flanks (KL < cutoff) are cut on both sides
    rng = np.random.default_rng(3)
    for _ in range(40):
        m1, m2 = synth.synth_pwms(2, int(rng.integers(0, 10 ** 6)), nmin=3, nmax=12)
        w1, w2 = [int(x) for x in rng.integers(1, 5, m1.shape[0])], [int(x) for x in rng.integers(1, 5, m2.shape[0])]
        s = int(rng.integers(-m2.shape[0] + 1, m1.shape[0]))
        got, gw = M.mergePWMWeighted((Mo.PWM(None, m1), w1), (Mo.PWM(None, m2), w2), s)
        exp, ew = app_oracle.merge_pwm_weighted((m1, w1), (m2, w2), s)
        assert gw == ew and np.array_equal(got._mat, exp)
