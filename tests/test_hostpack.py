"""The host side of the packed genome store (csrc/hostpack.cu, no GPU needed): pwmscan_pack_host writes the planes the
device packer would write — 2 bits per base (16 bases per word, base t at bits 2t..2t+1, A C G T = 0 1 2 3), 1 mask bit per
base for anything else — and reports fromBS's verdict (Seq.hs:86-95)."""
import numpy as np
import pytest

from bioinformatics_toolkit_b200 import _lib, synth


def _reference(data, uppercase):
    b = np.frombuffer(bytes(data), dtype=np.uint8).copy()
    if uppercase:
        low = (b >= ord("a")) & (b <= ord("z"))
        b[low] -= 32
    code = np.full(b.size, 255, np.uint8)
    for c, ch in enumerate(b"ACGT"):
        code[b == ch] = c
    units = (b.size + 31) // 32
    valid = np.zeros(units * 32, bool)
    valid[:b.size] = code != 255
    c2 = np.zeros(units * 32, np.uint32)
    c2[:b.size] = np.where(code != 255, code, 0)
    mask = (~valid).reshape(units, 32)
    mask_words = (mask.astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(axis=1).astype(np.uint32)
    iupac = np.isin(b, np.frombuffer(b"ACGTNVHDBMKWSYR", dtype=np.uint8))
    bad_a = np.flatnonzero(code == 255)
    bad_i = np.flatnonzero(~iupac)
    return c2.reshape(-1, 16), valid.reshape(-1, 16), mask_words, (int(bad_a[0]) if bad_a.size else -1), (int(bad_i[0]) if bad_i.size else -1)


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 1000, 70001])
def test_pack_host_matches_the_layout(n):
    rng = np.random.default_rng(n)
    data = rng.choice(np.frombuffer(b"ACGTACGTACGTacgtNRYn", dtype=np.uint8), n)
    for uppercase in (False, True):
        p = _lib.PackedSeq(data, uppercase=uppercase)
        c2, valid, mask_words, bad_a, bad_i = _reference(data, uppercase)
        assert p.seq2.size == 2 * ((n + 31) // 32) and p.mask.size == (n + 31) // 32
        got = (p.seq2[:, None] >> (2 * np.arange(16, dtype=np.uint32))) & 3
        assert np.array_equal(got[valid], c2[valid])                     # codes of the A/C/G/T bases (the others carry a hash)
        assert np.array_equal(p.mask, mask_words)
        assert (p.first_non_acgt, p.first_non_iupac) == (bad_a, bad_i)


def test_pack_host_threads_agree_and_invalid_bytes_are_reported():
    dna = synth.synth_dna(3_000_000, 9, n_runs=[(1_000_000, 5000)])
    one = _lib.PackedSeq(dna, nthreads=1)
    many = _lib.PackedSeq(dna, nthreads=8)
    assert np.array_equal(one.seq2, many.seq2) and np.array_equal(one.mask, many.mask)
    assert one.first_non_acgt == many.first_non_acgt == 1_000_000 and one.first_non_iupac == -1
    bad = dna.copy()
    bad[2_500_000] = ord("!")
    assert _lib.PackedSeq(bad, nthreads=8).first_non_iupac == 2_500_000
    s2, mk, length = one.slice(64, 1000)
    assert length == 936 and s2[0] == one.seq2[4] and mk[0] == one.mask[2]
    with pytest.raises(ValueError):
        one.slice(33, 100)


def test_packed_sidecar_of_the_genome_index_roundtrip(tmp_path):
    """seq_io.mkPackedIndex / PackedGenome: the `.packed` file next to the index holds, per chromosome, exactly the planes
    pwmscan_pack_host makes of its upper-cased bytes, plus fromBS's verdict; slices are views into the mapped file."""
    from bioinformatics_toolkit_b200 import seq_io
    chroms = [(b"chr1", synth.synth_dna(100_003, 1, n_runs=[(5000, 100)])), (b"chr2", synth.synth_dna(33, 2)), (b"chrE", np.zeros(0, np.uint8)),
              (b"chrW", synth.synth_dna(5000, 3))]
    chroms[0][1][7:20] |= 0x20                       # lower case
    chroms[3][1][1234] = ord("!")                    # outside the IUPAC alphabet
    path = str(tmp_path / "g.idx")
    seq_io.write_genome(chroms, path)
    with seq_io.Genome(path) as g:
        pth = seq_io.mkPackedIndex(g)
    assert pth == path + ".packed"
    pg = seq_io.PackedGenome(pth)
    for nm, arr in chroms:
        pl, ref = pg.planes(nm), _lib.PackedSeq(arr, uppercase=True)
        assert pl.length == arr.size and np.array_equal(pl.seq2, ref.seq2) and np.array_equal(pl.mask, ref.mask)
        assert (pl.first_non_acgt, pl.first_non_iupac) == (ref.first_non_acgt, ref.first_non_iupac)
    assert pg.planes(b"chr1").first_non_acgt == 5000 and pg.planes(b"chrW").first_non_iupac == 1234 and pg.planes("chr2").first_non_acgt == -1
    s2, mk, n = pg.planes(b"chr1").slice(64, 1000)
    assert n == 936 and s2[0] == pg.planes(b"chr1").seq2[4]
    with pytest.raises(ValueError):
        seq_io.PackedGenome(path)                    # the index file itself is not a packed store
