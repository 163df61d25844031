"""scan_motif / merge path through the Python mirror of the reference API (device compute) against
the oracle-side restatement of the host logic (oracle/app_oracle.py)."""
import io
import os

import numpy as np
import pytest

from bioinformatics_toolkit_b200 import merge as M
from bioinformatics_toolkit_b200 import motif as Mo
from bioinformatics_toolkit_b200 import synth
from bioinformatics_toolkit_b200.alignment import AlignFn, alignment
from bioinformatics_toolkit_b200.bed_utils import mkCutoffMotifs, scanGenome, scanMotif, streamBed3
from bioinformatics_toolkit_b200.seq_io import getChrSizes, getSeq, openGenome, write_genome

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small_genome(tmp_path_factory):
    d = tmp_path_factory.mktemp("genome")
    chroms = [(b"chr1", synth.synth_dna(60000, 31, gc=0.45, n_runs=[(20000, 900)], lower_frac=0.2)),
              (b"chr2", synth.synth_dna(25000, 32, gc=0.40)),
              (b"chrW", np.frombuffer(b"ACGTRYKMACGTNNNNACGTXACGT" * 40, dtype=np.uint8))]
    path = os.path.join(d, "genome.idx")
    write_genome(chroms, path)
    return path, dict(chroms)


def test_genome_index_roundtrip(small_genome):
    path, chroms = small_genome
    g = openGenome(path)
    assert getChrSizes(g) == [(b"chr1", 60000), (b"chr2", 25000), (b"chrW", 1000)]
    ok, dna = getSeq(g, (b"chr2", 100, 160))
    assert ok and dna.bs == bytes(chroms[b"chr2"][100:160]).upper()
    assert getSeq(g, (b"chr2", 100, 25001))[0] is False and getSeq(g, (b"chrQ", 0, 1))[0] is False
    assert getSeq(g, (b"chrW", 0, 30))[0] is False          # 'X' is not in the IUPAC alphabet


def test_scan_motif_app_lines(small_genome, tmp_path):
    from oracle import app_oracle
    path, chroms = small_genome
    rng = np.random.default_rng(5)
    mats = synth.synth_pwms(12, 44, nmin=6, nmax=18)
    names = [b"M%03d some name" % i for i in range(len(mats))]
    motifs = [Mo.Motif(nm, Mo.PWM(None, m)) for nm, m in zip(names, mats)]
    beds = []
    for _ in range(60):
        c = [b"chr1", b"chr2", b"chrW", b"chrMissing"][int(rng.integers(0, 4))]
        L = len(chroms.get(c, b"x" * 30000))
        s = int(rng.integers(0, max(1, L - 600)))
        beds.append((c, s, s + int(rng.integers(5, 700))))
    beds.append((b"chr2", 24900, 25100))                      # beyond the end: skipped
    bedfile = os.path.join(tmp_path, "in.bed")
    with open(bedfile, "wb") as f:
        for c, s, e in beds:
            f.write(b"%s\t%d\t%d\textra\t0\t+\n" % (c, s, e))
    p = 1e-3
    cms = mkCutoffMotifs(Mo.default_bkgd, p, motifs)
    out = io.BytesIO()
    n = scanMotif(openGenome(path), cms, streamBed3(bedfile), out=out, warn=None)
    got = out.getvalue().split(b"\n")[:-1] if n else []
    exp = app_oracle.scan_motif_lines(chroms, names, mats, beds, p=p)
    assert len(exp) > 50
    assert got == exp


def test_scan_genome_whole_chromosomes_from_the_store(small_genome):
    """scanGenome = scanMotif over [(chr, 0, size)] for every chromosome of the index file, read through pinned staging
    buffers in small chunks (several chunks + halos per chromosome, pwmscan_scan_host_many): BED6 text byte-identical to
    the oracle app on whole-chromosome regions; chrW holds a byte outside the IUPAC alphabet and is skipped (fromBS)."""
    from oracle import app_oracle
    path, chroms = small_genome
    mats = synth.synth_pwms(10, 45, nmin=6, nmax=19)
    names = [b"G%02d" % i for i in range(len(mats))]
    motifs = [Mo.Motif(nm, Mo.PWM(None, m)) for nm, m in zip(names, mats)]
    p = 1e-3
    cms = mkCutoffMotifs(Mo.default_bkgd, p, motifs)
    g = openGenome(path)
    beds = [(c, 0, n) for c, n in getChrSizes(g)]
    exp = app_oracle.scan_motif_lines(chroms, names, mats, beds, p=p)
    assert len(exp) > 200
    for chunk_len in (7000, 1 << 20):
        out = io.BytesIO()
        n = scanGenome(g, cms, out=out, chunk_len=chunk_len, warn=None)
        got = out.getvalue().split(b"\n")[:-1] if n else []
        assert got == exp, chunk_len
    tuples = scanGenome(g, cms, chroms=[b"chr2"], chunk_len=9000, warn=None)
    assert len(tuples) == sum(1 for l in exp if l.startswith(b"chr2\t"))
    # the same from the packed sidecar of the index (mkindex --packed): slices of the 2-bit + mask planes instead of bytes
    from bioinformatics_toolkit_b200.seq_io import PackedGenome, mkPackedIndex
    pg = PackedGenome(mkPackedIndex(g))
    for chunk_len in (7000, 1 << 20):
        out = io.BytesIO()
        n = scanGenome(g, cms, out=out, chunk_len=chunk_len, warn=None, packed=pg)
        got = out.getvalue().split(b"\n")[:-1] if n else []
        assert got == exp, ("packed", chunk_len)


def test_iterative_merge_and_tree(oracle):
    from oracle import app_oracle
    mats = synth.synth_pwms(14, 71, nmin=6, nmax=14)
    mats += [oracle.rc_pwm(mats[2]), mats[4].copy(), 0.5 * (mats[6] + 0.25)]
    names = [b"m%d" % i for i in range(len(mats))]
    motifs = [Mo.Motif(nm, Mo.PWM(None, m)) for nm, m in zip(names, mats)]
    got = M.iterativeMerge(alignment, 0.2, motifs)
    exp = app_oracle.iterative_merge(0.2, names, mats)
    assert [g[0] for g in got] == [e[0] for e in exp]
    assert len(got) < len(mats)
    for g, e in zip(got, exp):
        assert g[2] == e[2] and np.allclose(g[1]._mat, e[1], rtol=0, atol=1e-15)
    tree = M.buildTree(AlignFn("exp", 0.05, "l1"), motifs)
    assert sorted(m._name for m in M.flatten(tree)) == sorted(names)
    cl = M.cutAt(tree, 0.2)
    assert 1 < len(cl) < len(mats)
    merged = M.dilute(M.mergeTreeWeighted(alignment, cl[0]))
    assert merged._mat.shape[1] == 4 and np.allclose(merged._mat.sum(axis=1), 1.0, atol=1e-9)


def test_reference_api_invariants(ref_motifs):
    """The five invariants of the reference's own test-suite (Tests/Motif.hs:58-97), through the
    drop-in API with the device doing the work."""
    from bioinformatics_toolkit_b200 import search as S
    from bioinformatics_toolkit_b200.seq import DNA, rc
    bg = Mo.default_bkgd
    dna = DNA(bytes(synth.synth_dna(5000, 2)), "Basic")
    pwms = [Mo.PWM(None, m) for _, m in ref_motifs]
    th = 0.6 * Mo.optimalScore(bg, pwms[0])
    assert S.findTFBS(bg, pwms[0], dna, th, True) == S.findTFBSSlow(bg, pwms[0], dna, th)
    for p in pwms:
        assert S.maxMatchSc(bg, p, dna) == Mo.scores(bg, p, dna).max()
        c = Mo.scoreCDF(bg, p)
        assert Mo.cdf_(c, 1 - 1e-4) == Mo.cdf_(Mo.truncateCDF(0.999, c), 1 - 1e-4) == Mo.pValueToScore(1e-4, bg, p)
        a = np.round(1e5 * Mo.scores(bg, p, dna))
        b = np.round(1e5 * Mo.scores(bg, Mo.rcPWM(p), rc(dna))[::-1])
        assert np.array_equal(a, b)


def test_command_line_apps(small_genome, tmp_path, oracle):
    """apps/scan_motif.py and apps/merge_motifs.py end to end (argument parsing, file formats)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path, chroms = small_genome
    mats = synth.synth_pwms(6, 101, nmin=8, nmax=14)
    motifs = [Mo.Motif(b"motif_%d" % i, Mo.PWM(20, m)) for i, m in enumerate(mats)]
    meme = os.path.join(tmp_path, "m.meme")
    Mo.writeMEME(meme, motifs)
    back = Mo.readMEME(meme)
    assert [m._name for m in back] == [m._name for m in motifs]
    assert all(np.all(np.abs(a._pwm._mat - b._pwm._mat) <= np.spacing(b._pwm._mat)) for a, b in zip(back, motifs))   # read_double: exact or one ulp
    mats = [m._pwm._mat for m in back]                      # what the apps see: the file's text through readDouble
    bed = os.path.join(tmp_path, "r.bed")
    beds = [(b"chr1", 100, 5100), (b"chr2", 0, 3000), (b"chr9", 1, 2)]
    with open(bed, "wb") as f:
        for c, s, e in beds:
            f.write(b"%s\t%d\t%d\n" % (c, s, e))
    r = subprocess.run([sys.executable, os.path.join(root, "apps", "scan_motif.py"), path, meme, bed, "-p", "1e-3"],
                       capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    from oracle import app_oracle
    exp = app_oracle.scan_motif_lines(chroms, [m._name for m in motifs], mats, beds, p=1e-3)
    assert r.stdout.split(b"\n")[:-1] == exp and len(exp) > 10
    assert b"Warning: no sequence for region" in r.stderr
    # merge_motifs dist
    r = subprocess.run([sys.executable, os.path.join(root, "apps", "merge_motifs.py"), "dist", "-a", meme], capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    lines = r.stdout.split(b"\n")[:-1]
    assert len(lines) == 15
    a, b, d = lines[0].split(b"\t")
    od, _ = oracle.alignment(mats[0], mats[1])
    assert (a, b) == (b"motif_0", b"motif_1") and abs(float(d) - od) <= 1e-12 * max(od, 1e-300)
    # merge_motifs merge (iter) and cluster
    outp = os.path.join(tmp_path, "merged.meme")
    r = subprocess.run([sys.executable, os.path.join(root, "apps", "merge_motifs.py"), "merge", meme, "-t", "0.3", "-o", outp], capture_output=True, timeout=300)
    assert r.returncode == 0 and b"Read 6 motifs" in r.stderr, r.stderr.decode()[-2000:]
    assert 1 <= len(Mo.readMEME(outp)) <= 6
    r = subprocess.run([sys.executable, os.path.join(root, "apps", "merge_motifs.py"), "cluster", meme, "-h", "0.3"], capture_output=True, timeout=300)
    assert r.returncode == 0 and sorted(b"\t".join(r.stdout.split(b"\n")[:-1]).split(b"\t")) == sorted(m._name for m in motifs)


def test_device_affinity_equals_host_formula():
    """pwmscan_hits_affinity: per-hit p = 1 - cdf score (bit-identical fp64) and toAffinity p (identical
    integers, including values near a rounding tie) against the host restatement of Utils.hs:109-123."""
    from bioinformatics_toolkit_b200 import _lib, bed_utils
    ctx = _lib.default_context(0)
    mats = synth.synth_pwms(40, 91, nmin=6, nmax=22)
    motifs = [Mo.Motif(b"m%d" % i, Mo.PWM(None, m)) for i, m in enumerate(mats)]
    cms = mkCutoffMotifs(Mo.default_bkgd, 1e-3, motifs, ctx)
    mo = _lib.DeviceMotifs(ctx, mats, Mo.default_bkgd.freqs)
    plan = _lib.Plan(mo, [c._cutoff for c in cms])
    h = plan.scan_host(synth.synth_dna(400_000, 17, gc=0.47))
    assert len(h) > 5000
    off = np.concatenate([[0], np.cumsum([c._cdf.xs.size for c in cms])]).astype(np.int64)
    cx = np.concatenate([c._cdf.xs for c in cms])
    cy = np.concatenate([c._cdf.ps for c in cms])
    aff, pv = h.affinity(off, cx, cy, want_pvalue=True)
    exp_p = np.empty(len(h))
    for m in range(len(cms)):
        k = h.motif == m
        exp_p[k] = 1 - bed_utils._cdf_vec(cms[m]._cdf, h.score[k])
    assert np.array_equal(pv.view(np.uint64), exp_p.view(np.uint64))
    assert np.array_equal(aff, np.array([bed_utils.toAffinity(float(p)) for p in exp_p], dtype=np.int64))
    assert aff.min() >= 0 and aff.max() <= 1000 and np.unique(aff).size > 20
    # scores outside the CDF's support and exactly on its points
    aff2 = h.affinity(off, cx + 1000.0, cy)            # every score left of the first point: cdf 0 -> p = 1 -> affinity of 1.0
    assert np.all(aff2 == bed_utils.toAffinity(1.0))
    aff3 = h.affinity(off, cx - 1000.0, cy)            # every score right of the last point: cdf = last p
    last_p = np.array([1 - c._cdf.ps[-1] for c in cms])
    assert np.array_equal(aff3, np.array([bed_utils.toAffinity(float(last_p[m])) for m in h.motif], dtype=np.int64))


def test_iterative_merge_resident_equals_per_call_path():
    """iterativeMerge with the PWMs resident on the device (pwmscan_alnset_*: slot updates, row refreshes, per-row
    minimum tracking) gives exactly the result of the per-call path (motif set rebuilt and the full matrix re-scanned
    at every merge) — 160 motifs with duplicates, reverse-complement duplicates and palindromes, two thresholds."""
    mats = synth.config_c5_motifs(160, seed=41)
    motifs = [Mo.Motif(b"m%d" % i, Mo.PWM(20, m)) for i, m in enumerate(mats)]

    class PerCall:                       # same device alignments, but without `.resident`
        def __init__(self, f):
            self.all_pairs, self.pairs = f.all_pairs, f.pairs

    for th in (0.1, 0.3):
        a = M.iterativeMerge(alignment, th, motifs)
        b = M.iterativeMerge(PerCall(alignment), th, motifs)
        assert len(a) == len(b) and len(a) < len(motifs)
        for x, y in zip(a, b):
            assert x[0] == y[0] and x[2] == y[2] and np.array_equal(x[1]._mat, y[1]._mat)


def test_scan_motif_config_c3_sample_at_size(tmp_path):
    """BASELINE config 3 shape at a size the oracle finishes in seconds: the 800 JASPAR-sized PWMs x 2 000 BED regions of
    500 bp (unsorted, on four chromosomes, some soft-masked, some out of range) at p = 1e-5 — scan_motif's BED6 text must
    be byte-identical to the oracle composition (per-region multi-threaded oracle scan + cdf + toAffinity + toLine)."""
    from oracle import app_oracle
    chroms = [(b"chr1", synth.synth_dna(3_000_000, 4, gc=0.41, lower_frac=0.3, n_runs=[(1_350_000, 90_000)])),
              (b"chr2", synth.synth_dna(2_000_000, 104, gc=0.41)),
              (b"chrX", synth.synth_dna(1_200_000, 2204, gc=0.38, lower_frac=0.1)),
              (b"chrM", synth.synth_dna(16_569, 2404, gc=0.44))]
    path = os.path.join(tmp_path, "genome.idx")
    write_genome(chroms, path)
    gen = dict(chroms)
    mats = synth.config_c4_motifs(800)
    names = [b"MA%04d.1" % i for i in range(len(mats))]
    motifs = [Mo.Motif(nm, Mo.PWM(None, m)) for nm, m in zip(names, mats)]
    rng = np.random.default_rng(3)
    sizes = np.array([c[1].size for c in chroms], dtype=np.float64)
    beds = []
    for _ in range(2000):
        ci = int(rng.choice(len(chroms), p=sizes / sizes.sum()))
        s = int(rng.integers(0, chroms[ci][1].size - 400))            # a few run past the end: skipped with a warning
        beds.append((chroms[ci][0], s, s + 500))
    cms = mkCutoffMotifs(Mo.default_bkgd, 1e-5, motifs)
    out = io.BytesIO()
    n = scanMotif(openGenome(path), cms, iter(beds), out=out, warn=None)
    got = out.getvalue().split(b"\n")[:-1]
    exp = app_oracle.scan_motif_lines_mt(gen, names, mats, beds, p=1e-5, nthreads=os.cpu_count() or 8)
    assert n == len(got) and len(exp) > 20_000
    assert got == exp


def test_iterative_merge_device_matrix_and_long_merges():
    """pwmscan_alnset_matrix_init / _merge_step (matrix, row refresh and argmin resident on the device) against (a) the
    host-matrix path over the same resident PWMs, (b) the oracle's restatement of Merge.hs:113-170 — including a chain of
    merges whose PWM grows past 64 columns (the slots of the alignment set grow; the reference has no length limit)."""
    from oracle import app_oracle

    class HostMatrix:                    # resident PWMs, but the matrix and the argmin stay in Python (round-1 path)
        def __init__(self, f):
            self.f = f
            self.all_pairs, self.pairs = f.all_pairs, f.pairs

        def resident(self, pwms):
            r = self.f.resident(pwms)

            class R:
                pairs, update, free = r.pairs, r.update, r.free
            return R()

    mats = synth.config_c5_motifs(120, seed=43)
    motifs = [Mo.Motif(b"m%d" % i, Mo.PWM(20, m)) for i, m in enumerate(mats)]
    for th in (0.05, 0.25):
        a = M.iterativeMerge(alignment, th, motifs)
        b = M.iterativeMerge(HostMatrix(alignment), th, motifs)
        assert len(a) == len(b) and len(a) < len(motifs)
        for x, y in zip(a, b):
            assert x[0] == y[0] and x[2] == y[2] and np.array_equal(x[1]._mat, y[1]._mat)
    # overlapping windows of one long PWM merge back into something longer than 64 columns
    master = synth.synth_pwm(150, np.random.Generator(np.random.PCG64(5)), alpha=0.05, zero_frac=0.0)
    tiles = [np.ascontiguousarray(master[s:s + 24]) for s in range(0, 126, 6)]
    tmot = [Mo.Motif(b"t%d" % i, Mo.PWM(10, m)) for i, m in enumerate(tiles)]
    lin = AlignFn("linear", 0.001, "l1")
    got = M.iterativeMerge(lin, 0.05, tmot)
    assert max(x[1]._mat.shape[0] for x in got) > 128 and len(got) < len(tmot)
    exp = app_oracle.iterative_merge(0.05, [m._name for m in tmot], tiles, penal="linear", gap=0.001, comb="l1")
    assert len(got) == len(exp)
    for x, y in zip(got, exp):
        assert x[0] == y[0] and list(x[2]) == list(y[2]) and np.array_equal(x[1]._mat, y[1])
