"""Parity of the CUDA scan (through the C ABI) with the oracle: bit-identical hit sets
(motif, strand, position) AND bit-identical fp64 scores."""
import json
import os

import numpy as np
import pytest

from bioinformatics_toolkit_b200 import _lib, synth

pytestmark = pytest.mark.gpu
BG = (0.25, 0.25, 0.25, 0.25)
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def gpu_scan(ctx, mats, thres, dna, strands=2, skip=True, bg=BG, seg_off=None, keep_ascii=False):
    seq = _lib.DeviceSeq(ctx, dna, seg_off=seg_off, keep_ascii=keep_ascii or not skip)
    mo = _lib.DeviceMotifs(ctx, mats, bg)
    return _lib.Plan(mo, thres, strands=strands, skip_ambiguous=skip).scan(seq)


def oracle_scan(oracle, mats, thres, dna, strands=2, bg=BG, skip=True):
    cols = []
    for m, (mat, th) in enumerate(zip(mats, thres)):
        for s in range(strands):
            mm = mat if s == 0 else oracle.rc_pwm(mat)
            pos, sc = oracle.find_tfbs(bg, mm, dna, th, skip, mode=1)
            cols.append((np.full(pos.size, m, np.int32), np.full(pos.size, s, np.uint8), pos, sc))
    return tuple(np.concatenate([c[k] for c in cols]) for k in range(4))


def assert_same(h, o):
    assert len(h) == o[0].size
    assert np.array_equal(h.motif, o[0]) and np.array_equal(h.strand, o[1]) and np.array_equal(h.pos, o[2])
    assert np.array_equal(h.score.view(np.uint64), o[3].view(np.uint64))     # bit-identical fp64


def test_mixed_lengths_both_strands(ctx, oracle):
    mats = synth.synth_pwms(150, 21, nmin=6, nmax=24)
    dna = synth.synth_dna(400000, 9, gc=0.45, n_runs=[(0, 40), (30000, 700), (399990, 10)])
    thres = [0.55 * oracle.optimal_score(BG, m) for m in mats]
    h = gpu_scan(ctx, mats, thres, dna)
    cnt, o = oracle.scan_mt(BG, mats, thres, dna, mode=1, nthreads=8, cap=1 << 23)
    assert cnt > 1000
    assert_same(h, o)


def test_reference_fixture_goldens(ctx, oracle, ref_motifs):
    g = json.load(open(os.path.join(GOLDEN, "oracle_goldens.json")))
    dna = synth.synth_dna(5000, 2)
    mats = [m for _, m in ref_motifs]
    thres = [0.6 * oracle.optimal_score(BG, m) for m in mats]
    h = gpu_scan(ctx, mats, thres, dna, strands=1)
    for m, e in enumerate(g["fasta"]):
        hm = h.select(motif=m)
        assert hm.pos.tolist() == e["tfbs_pos"]
        assert [float(v).hex() for v in hm.score] == e["tfbs_sc"]


def test_long_motifs_and_edge_cutoffs(ctx, oracle):
    bg = (0.3, 0.2, 0.2, 0.3)
    mats = synth.synth_pwms(6, 5, nmin=29, nmax=60) + synth.synth_pwms(9, 8, nmin=5, nmax=16)
    dna = synth.synth_dna(30000, 6, gc=0.4, n_runs=[(5000, 64)])
    opt = [max(oracle.optimal_score(bg, m), 1.0) for m in mats]
    thres = [0.3 * o for o in opt]
    thres[7] = 1.5 * opt[7]
    thres[8] = -1e9
    thres[9] = float("inf")
    thres[10] = float("-inf")
    thres[11] = 0.0
    h = gpu_scan(ctx, mats, thres, dna[:4000], bg=bg)
    o = oracle_scan(oracle, mats, thres, dna[:4000], bg=bg)
    assert_same(h, o)
    assert len(h.select(motif=7)) == 0 and len(h.select(motif=9)) == 0
    assert len(h.select(motif=8, strand=0)) == 4000 - mats[8].shape[0] + 1


def test_segments_never_span(ctx, oracle):
    mats = synth.synth_pwms(20, 33, nmin=6, nmax=20)
    lens = [0, 5, 500, 37, 1200, 19, 20, 3000, 1]
    segs = [synth.synth_dna(L, 50 + i, n_runs=[(L // 3, 3)] if L > 100 else ()) for i, L in enumerate(lens)]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    thres = [0.45 * oracle.optimal_score(BG, m) for m in mats]
    h = gpu_scan(ctx, mats, thres, np.concatenate(segs), seg_off=off)
    tot = 0
    for s, d in enumerate(segs):
        o = oracle_scan(oracle, mats, thres, d)
        hs = h.select(segment=s)
        assert_same(hs, o)
        tot += o[0].size
    assert tot == len(h) and tot > 100
    assert np.all(np.diff(h.segment) >= 0)


def test_iupac_aware_search(ctx, oracle, ref_motifs):
    mats = [m for _, m in ref_motifs]
    d = bytearray(bytes(synth.synth_dna(20000, 5)))
    for i, ch in enumerate(b"NVHDBMKWSYR"):
        d[1000 + 37 * i] = ch
    d[5000:5040] = b"N" * 40
    thres = [0.4 * oracle.optimal_score(BG, m) for m in mats]
    h = gpu_scan(ctx, mats, thres, bytes(d), skip=False)
    o = oracle_scan(oracle, mats, thres, bytes(d), skip=False)
    assert_same(h, o)
    d[777] = ord("X")
    with pytest.raises(_lib.InvalidNucleotide):
        gpu_scan(ctx, mats, [-1e9] * len(mats), bytes(d), skip=False)


def test_lowercase_and_uppercase_flag(ctx, oracle, ref_motifs):
    mats = [m for _, m in ref_motifs]
    raw = synth.synth_dna(50000, 12, lower_frac=0.3)
    thres = [0.5 * oracle.optimal_score(BG, m) for m in mats]
    mo = _lib.DeviceMotifs(ctx, mats, BG)
    plan = _lib.Plan(mo, thres)
    # raw bytes: lower case is not ACGT (lookAheadSearch' scores it -inf)
    assert_same(plan.scan(_lib.DeviceSeq(ctx, raw)), oracle_scan(oracle, mats, thres, raw))
    # fromBS semantics: upper-cased first
    up = np.frombuffer(bytes(raw).upper(), dtype=np.uint8)
    s = _lib.DeviceSeq(ctx, raw, uppercase=True)
    assert_same(plan.scan(s), oracle_scan(oracle, mats, thres, up))
    assert s.first_invalid("Basic") == -1
    assert _lib.DeviceSeq(ctx, raw).first_invalid("Basic") == int(np.nonzero(raw >= 97)[0][0])


def test_tiny_and_empty(ctx, oracle, ref_motifs):
    mats = [m for _, m in ref_motifs]
    for dna in (b"", b"ACGT", b"ACGTACGTAC", bytes(synth.synth_dna(11, 1)), bytes(synth.synth_dna(33, 1))):
        h = gpu_scan(ctx, mats, [0.0] * len(mats), dna)
        o = oracle_scan(oracle, mats, [0.0] * len(mats), np.frombuffer(dna, dtype=np.uint8))
        assert_same(h, o)


def test_dense_scores_and_max(ctx, oracle, ref_motifs, ref_motifs_meme):
    mats = [m for _, m in ref_motifs] + [m for _, m in ref_motifs_meme]
    d = bytearray(bytes(synth.synth_dna(70000, 15)))
    for i, ch in enumerate(b"NVHDBMKWSYR"):
        d[2000 + 41 * i] = ch
    seq = _lib.DeviceSeq(ctx, bytes(d), keep_ascii=True)
    mo = _lib.DeviceMotifs(ctx, mats, BG)
    mx = _lib.max_match_sc(seq, mo)
    for m, mat in enumerate(mats):
        for s in (0, 1):
            ref = oracle.scores(BG, mat if s == 0 else oracle.rc_pwm(mat), bytes(d))
            got = _lib.scores(seq, mo, m, s)
            assert np.array_equal(got.view(np.uint64), ref.view(np.uint64))
        assert mx[m] == oracle.max_match_sc(BG, mat, bytes(d))
    d[100] = ord("!")
    with pytest.raises(_lib.InvalidNucleotide):
        _lib.scores(_lib.DeviceSeq(ctx, bytes(d), keep_ascii=True), mo, 0, 0)
    short = _lib.DeviceSeq(ctx, b"ACG", keep_ascii=True)
    assert _lib.scores(short, mo, 0, 0).size == 0 and np.all(np.isneginf(_lib.max_match_sc(short, mo)))


def test_config_c1_full_size(ctx, oracle):
    """BASELINE config 1: one CTCF-like 19-bp PWM over 10 Mbp, both strands, vs the oracle."""
    dna, mats = synth.config_c1()
    th = [json.load(open(os.path.join(GOLDEN, "config_thresholds.json")))["c1"][0]]
    th = [float.fromhex(th[0])]
    up = np.frombuffer(bytes(dna).upper(), dtype=np.uint8)
    h = _lib.Plan(_lib.DeviceMotifs(ctx, mats, BG), th).scan(_lib.DeviceSeq(ctx, dna, uppercase=True))
    cnt, o = oracle.scan_mt(BG, mats, th, up, mode=1, nthreads=8, cap=1 << 20)
    assert 20 < cnt < 5000
    assert_same(h, o)


def test_config_c2_full_size_properties(ctx, oracle):
    """BASELINE config 2 (100 PWMs x 248 956 422 bp): size-independent properties + an oracle slice."""
    dna, mats = synth.config_c2()
    th = [float.fromhex(v) for v in json.load(open(os.path.join(GOLDEN, "config_thresholds.json")))["c2"]]
    seq = _lib.DeviceSeq(ctx, dna)
    mo = _lib.DeviceMotifs(ctx, mats, BG)
    h = _lib.Plan(mo, th).scan(seq)
    assert 100000 < len(h) < 3000000
    # sorted (motif, strand, pos), no duplicates
    key = (h.motif.astype(np.int64) << 40) | (h.strand.astype(np.int64) << 39) | h.pos
    assert np.all(np.diff(key) > 0)
    # idempotence / determinism
    h2 = _lib.Plan(mo, th).scan(seq)
    assert np.array_equal(h.pos, h2.pos) and np.array_equal(h.score.view(np.uint64), h2.score.view(np.uint64))
    # strand symmetry: strand-1 hits of motif m == strand-0 hits of rcPWM(m), same cutoff
    rcs = [oracle.rc_pwm(m) for m in mats[:8]]
    hr = _lib.Plan(_lib.DeviceMotifs(ctx, rcs, BG), th[:8], strands=1).scan(seq)
    for m in range(8):
        a, b = h.select(motif=m, strand=1), hr.select(motif=m)
        assert np.array_equal(a.pos, b.pos) and np.array_equal(a.score.view(np.uint64), b.score.view(np.uint64))
    # no hit inside the N blocks
    for (p, ln) in synth.chr1_n_runs():
        inside = (h.pos + 7 >= p) & (h.pos < p + ln - 30)
        assert not inside.any()
    # oracle on two slices (one straddling the 18 Mbp N block's end)
    for lo in (5_000_000, 139_699_000):
        hi = lo + 1_000_000
        cnt, o = oracle.scan_mt(BG, mats, th, dna[lo:hi], mode=1, nthreads=8, cap=1 << 22)
        nlen = np.array([m.shape[0] for m in mats])[h.motif]
        k = (h.pos >= lo) & (h.pos + nlen <= hi)
        assert k.sum() == cnt
        assert np.array_equal(h.motif[k], o[0]) and np.array_equal(h.strand[k], o[1]) and np.array_equal(h.pos[k] - lo, o[2])
        assert np.array_equal(h.score[k].view(np.uint64), o[3].view(np.uint64))


def test_scan_host_pipeline_equals_resident_scan(ctx, oracle):
    """pwmscan_scan_host (chunked upload overlapped with the scan) == seq_upload + scan, including
    sequences spanning several 16 Mi-base chunks, segments and the IUPAC-aware path."""
    mats = synth.synth_pwms(70, 3, nmin=6, nmax=22)
    thres = [0.62 * oracle.optimal_score(BG, m) for m in mats]
    mo = _lib.DeviceMotifs(ctx, mats, BG)
    plan = _lib.Plan(mo, thres)
    for L in (1000, 16_777_216 - 5, 40_000_000):
        dna = synth.synth_dna(L, 77, gc=0.45, n_runs=[(L // 2, 2000)], lower_frac=0.05)
        a = plan.scan_host(dna, uppercase=True)
        b = plan.scan(_lib.DeviceSeq(ctx, dna, uppercase=True))
        assert len(a) == len(b) and len(a) > 0
        for x, y in ((a.motif, b.motif), (a.strand, b.strand), (a.pos, b.pos), (a.score.view(np.uint64), b.score.view(np.uint64))):
            assert np.array_equal(x, y)
    lens = [300, 0, 5000, 17]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    dna = synth.synth_dna(int(off[-1]), 5)
    a = plan.scan_host(dna, seg_off=off)
    b = plan.scan(_lib.DeviceSeq(ctx, dna, seg_off=off))
    assert np.array_equal(a.pos, b.pos) and np.array_equal(a.segment, b.segment)
    p2 = _lib.Plan(mo, thres, skip_ambiguous=False)
    d = bytearray(bytes(synth.synth_dna(3000, 8)))
    d[100] = ord("N")
    o = oracle_scan(oracle, mats, thres, bytes(d), skip=False)
    assert_same(p2.scan_host(bytes(d)), o)


def test_single_motif_replicated_lanes(ctx, oracle):
    """Few motifs: the block's table columns are replicated and every lane replica scans its own
    stretch of the sequence; result must not change."""
    for count in (1, 2, 3, 9, 17):
        mats = synth.synth_pwms(count, 100 + count, nmin=6, nmax=20)
        dna = synth.synth_dna(2_000_000, 31 + count, gc=0.5, n_runs=[(1_000_000, 777)])
        thres = [0.6 * oracle.optimal_score(BG, m) for m in mats]
        h = gpu_scan(ctx, mats, thres, dna)
        cnt, o = oracle.scan_mt(BG, mats, thres, dna, mode=1, nthreads=8, cap=1 << 22)
        assert cnt > 0
        assert_same(h, o)


def test_repetitive_sequence_overflows_and_retries(ctx, oracle):
    """Low-complexity DNA makes far more candidates than the plan's iid estimate: the scan must grow
    its buffers and re-run (never truncate), resident and pipelined alike."""
    unit = b"ACGTTGCAAC"
    motif = np.full((10, 4), 0.01)
    for k, ch in enumerate(unit):
        motif[k, b"ACGT".index(ch)] = 0.97
    mats = [motif] + synth.synth_pwms(5, 9, nmin=8, nmax=12)
    L = 30_000_000
    dna = np.frombuffer(unit * (L // len(unit)), dtype=np.uint8).copy()
    dna[123_456:123_500] = ord("N")
    thres = [0.9 * oracle.optimal_score(BG, m) for m in mats]
    mo = _lib.DeviceMotifs(ctx, mats, BG)
    plan = _lib.Plan(mo, thres)
    h = plan.scan(_lib.DeviceSeq(ctx, dna))
    n0 = len(h.select(motif=0, strand=0))
    assert n0 >= L // len(unit) - 20                       # one forward hit per repeat unit
    assert len(h) > 2_500_000
    # oracle on a slice
    lo, hi = 100_000, 160_000
    cnt, o = oracle.scan_mt(BG, mats, thres, dna[lo:hi], mode=1, nthreads=4, cap=1 << 20)
    nlen = np.array([m.shape[0] for m in mats])[h.motif]
    k = (h.pos >= lo) & (h.pos + nlen <= hi)
    assert k.sum() == cnt and np.array_equal(h.pos[k] - lo, o[2]) and np.array_equal(h.score[k].view(np.uint64), o[3].view(np.uint64))
    h2 = plan.scan_host(dna)
    assert len(h2) == len(h) and np.array_equal(h2.pos, h.pos) and np.array_equal(h2.motif, h.motif)


def test_many_segments_and_many_motifs(ctx, oracle):
    rng = np.random.default_rng(4)
    lens = rng.integers(0, 60, size=40000)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    dna = synth.synth_dna(int(off[-1]), 19)
    mats = synth.synth_pwms(200, 55, nmin=5, nmax=30)
    thres = [0.7 * oracle.optimal_score(BG, m) for m in mats]
    h = gpu_scan(ctx, mats, thres, dna, seg_off=off)
    # reference: scan the concatenation, then drop windows that straddle a boundary
    cnt, o = oracle.scan_mt(BG, mats, thres, dna, mode=1, nthreads=8, cap=1 << 22)
    nlen = np.array([m.shape[0] for m in mats])[o[0]]
    seg = np.searchsorted(off, o[2], side="right") - 1
    inside = o[2] + nlen <= off[seg + 1]
    order = np.lexsort((o[2][inside], o[1][inside], o[0][inside], seg[inside]))
    assert len(h) == inside.sum() and len(h) > 1000
    assert np.array_equal(h.segment, seg[inside][order]) and np.array_equal(h.motif, o[0][inside][order])
    assert np.array_equal(h.pos, (o[2] - off[seg])[inside][order])
    assert np.array_equal(h.score.view(np.uint64), o[3][inside][order].view(np.uint64))


@pytest.mark.parametrize("variant", ["kmer", "kmer+bitmap", "kmer-bitmap", "lds", "mma"])
def test_prefilter_variant_parity(variant):
    """Every first stage — the k-mer mask tables in L2 (kmer; with the first-table bitmap forced on and off),
    the shared-memory 4-mer deficit tables (lds) and the experimental tcgen05 int8 one-hot x deficit GEMM
    (mma), selected with PWMSCAN_PREFILTER — must give the same bit-identical hit lists (own process: the
    mode is read once)."""
    import subprocess
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from bioinformatics_toolkit_b200 import _lib, synth
from oracle import oracle as O
BG = (0.25,) * 4
ctx = _lib.default_context(0)
for count, L, lo, hi in ((150, 300000, 6, 40), (3, 100000, 8, 12)):
    mats = synth.synth_pwms(count, 21 + count, nmin=lo, nmax=hi)
    dna = synth.synth_dna(L, 9, gc=0.45, n_runs=[(0, 40), (L // 3, 700)])
    thres = [0.55 * O.optimal_score(BG, m) for m in mats]
    h = _lib.Plan(_lib.DeviceMotifs(ctx, mats, BG), thres).scan_host(dna)
    cnt, o = O.scan_mt(BG, mats, thres, dna, mode=1, nthreads=8, cap=1 << 23)
    assert cnt == len(h) and cnt > 100
    assert np.array_equal(h.motif, o[0]) and np.array_equal(h.strand, o[1]) and np.array_equal(h.pos, o[2])
    assert np.array_equal(h.score.view(np.uint64), o[3].view(np.uint64))
assert ctx.timing()["prefilter_launches"] >= 1
print("VARIANT-PARITY-OK")
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PWMSCAN_PREFILTER=variant.split("+")[0].split("-")[0])
    if variant == "kmer+bitmap":
        env["PWMSCAN_FORCE_BITMAP"] = "1"        # every block through the two-phase bitmap path
    elif variant == "kmer-bitmap":
        env["PWMSCAN_FORCE_BITMAP"] = "0"        # every block through the plain gather path
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "VARIANT-PARITY-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_scan_workers_overlap_same_results(ctx):
    """shard.ScanWorkers: two contexts / host threads with scans in flight on one device give exactly the
    hit lists of a sequential scan, item by item, for resident and for host inputs."""
    from bioinformatics_toolkit_b200.shard import ScanWorkers
    mats = synth.synth_pwms(30, 77, nmin=6, nmax=24)
    mo = _lib.DeviceMotifs(ctx, mats, BG)
    th = mo.pvalue_to_score(1e-4)
    plan = _lib.Plan(mo, th)
    chunks = [synth.synth_dna(300_000 + 7919 * i, 100 + i, gc=0.4 + 0.02 * i, n_runs=[(1000 * i, 50)]) for i in range(5)]
    ref = [plan.scan_host(c) for c in chunks]
    w = ScanWorkers(0, mats, BG, th, workers=2)
    try:
        for hits, t in (w.scan(w.upload(chunks)), w.scan_host(chunks)):
            assert t["prefilter_launches"] >= len(chunks)
            for a, b in zip(hits, ref):
                assert len(a) == len(b) and len(b) > 50
                assert np.array_equal(a.motif, b.motif) and np.array_equal(a.strand, b.strand) and np.array_equal(a.pos, b.pos)
                assert np.array_equal(a.score.view(np.uint64), b.score.view(np.uint64))
    finally:
        w.close()


def test_kmer_tables_extreme_shapes(ctx, oracle):
    """The k-mer mask first stage at its edges: motifs of 1, 2, 7, 10, 11, 33, 63 and 64 columns in one plan
    (the tables see at most 32 columns, the middle stage all of them), forward-only with more than 256 motifs
    (two blocks of 256 single-strand combos), thresholds from loose to consensus-only, a short sequence that
    is mostly one work item, and a non-uniform background."""
    bg = (0.2, 0.3, 0.3, 0.2)
    lengths = [1, 2, 7, 10, 11, 33, 63, 64]
    mats = synth.synth_pwms(len(lengths), 31, lengths=lengths)
    dna = synth.synth_dna(70000, 12, gc=0.55, n_runs=[(0, 5), (33000, 200), (69900, 100)])
    opt = [oracle.optimal_score(bg, m) for m in mats]
    thres = [0.5 * o for o in opt]
    thres[3] = opt[3] - 1e-9          # consensus only
    thres[5], thres[6], thres[7] = -0.6 * opt[5], -1.0 * opt[6], -1.0 * opt[7]      # long motifs: cutoffs that random sequence reaches (71 / 258 / 403 hits)
    h = gpu_scan(ctx, mats, thres, dna[:20000], bg=bg)
    o = oracle_scan(oracle, mats, thres, dna[:20000], bg=bg)
    assert_same(h, o)
    assert len(h.select(motif=0)) > 1000 and len(h.select(motif=3)) < 10
    assert all(len(h.select(motif=m)) > 0 for m in (5, 6, 7)), [len(h.select(motif=m)) for m in range(8)]
    # forward-only, 300 motifs: blocks of 256 + 44 single-strand combos
    mats = synth.synth_pwms(300, 32, nmin=6, nmax=20)
    thres = [0.6 * oracle.optimal_score(BG, m) for m in mats]
    h = gpu_scan(ctx, mats, thres, dna, strands=1)
    assert_same(h, oracle_scan(oracle, mats, thres, dna, strands=1))
    assert len(h) > 1000


def test_random_sweep_small_cases(ctx, oracle):
    """Randomised sweep through the C ABI (fixed seeds): 1-8 motifs of 1..64 columns, skewed backgrounds, one or both
    strands, cutoffs from always-true to unreachable, short sequences with N runs — bit-identical to the oracle."""
    rng = np.random.default_rng(4242)
    for case in range(30):
        nm = int(rng.integers(1, 9))
        lengths = [int(rng.choice([1, 2, 3, 5, 8, 9, 10, 11, 16, 17, 21, 31, 33, 48, 64])) for _ in range(nm)]
        mats = synth.synth_pwms(nm, 700 + case, lengths=lengths, alpha=float(rng.choice([0.15, 0.3, 1.0])), zero_frac=float(rng.choice([0.0, 0.125, 0.4])))
        bgw = rng.dirichlet([8, 8, 8, 8])
        bg = tuple(float(x) for x in bgw / bgw.sum())
        dna = synth.synth_dna(int(rng.integers(1, 6000)), 1100 + case, gc=float(rng.uniform(0.3, 0.7)),
                              n_runs=[(int(rng.integers(0, 40)), int(rng.integers(1, 30)))] if rng.random() < 0.5 else [])
        thres = []
        for m in mats:
            opt = oracle.optimal_score(bg, m)
            thres.append(float(rng.choice([-1e9, -2.0 * abs(opt), -0.5 * abs(opt), 0.0, 0.3 * opt, 0.7 * opt, opt - 1e-9, opt, opt + 1.0])))
        strands = int(rng.integers(1, 3))
        h = gpu_scan(ctx, mats, thres, dna, strands=strands, bg=bg)
        o = oracle_scan(oracle, mats, thres, dna, strands=strands, bg=bg)
        assert_same(h, o)


def _same_hits(a, b):
    assert len(a) == len(b)
    for x, y in ((a.motif, b.motif), (a.strand, b.strand), (a.segment, b.segment), (a.pos, b.pos),
                 (a.score.view(np.uint64), b.score.view(np.uint64))):
        assert np.array_equal(x, y)


def test_scan_many_equals_single_scans(ctx, oracle):
    """pwmscan_scan_many / pwmscan_scan_host_many (several items in flight on their own lanes) hand back, in item
    order, exactly the hit lists of one pwmscan_scan per item — including empty items and more items than lanes."""
    mats = synth.synth_pwms(90, 5, nmin=6, nmax=21)
    thres = [0.58 * oracle.optimal_score(BG, m) for m in mats]
    plan = _lib.Plan(_lib.DeviceMotifs(ctx, mats, BG), thres)
    lens = [700_000, 0, 5, 1_300_000, 64, 2_000_000, 40_000, 900_000, 17_000_000]
    arrays = [synth.synth_dna(L, 300 + i, gc=0.42, n_runs=[(L // 2, 500)] if L > 1000 else ()) for i, L in enumerate(lens)]
    seqs = [_lib.DeviceSeq(ctx, a) for a in arrays]
    single = [plan.scan(s) for s in seqs]
    assert sum(len(h) for h in single) > 10000
    many = plan.scan_many(seqs)
    host = plan.scan_host_many(arrays)
    for a, b, c in zip(single, many, host):
        _same_hits(a, b)
        _same_hits(a, c)
    # streaming consumer: called in item order, results kept
    seen = []
    counts = plan.scan_many(seqs, consume=lambda h: (seen.append(len(h)), len(h))[1])
    assert counts == [len(h) for h in single] == seen
    assert ctx.timing()["hits"] == sum(counts) and ctx.timing()["kernel_launches"] >= 2 * len(seqs)
    # a consumer that raises does not leave the library in a bad state
    with pytest.raises(ZeroDivisionError):
        plan.scan_many(seqs, consume=lambda h: 1 // 0)
    _same_hits(single[3], plan.scan(seqs[3]))


def test_compact_columns_and_run_table(ctx, oracle):
    """The transferred form of a hit list: score + 32-bit positions + the (motif, strand) run table (single segment)
    or 16-bit motif|strand + segment (several segments); the wide columns are expanded from it."""
    mats = synth.synth_pwms(33, 8, nmin=6, nmax=18)
    thres = [0.55 * oracle.optimal_score(BG, m) for m in mats]
    dna = synth.synth_dna(600_000, 12, gc=0.47)
    h = gpu_scan(ctx, mats, thres, dna)
    ro = h.run_off
    assert ro is not None and ro.size == 2 * len(mats) + 1 and ro[0] == 0 and ro[-1] == len(h) and np.all(np.diff(ro) >= 0)
    for m in (0, 7, 32):
        for s in (0, 1):
            a, b = int(ro[2 * m + s]), int(ro[2 * m + s + 1])
            pos, sc = oracle.find_tfbs(BG, mats[m] if s == 0 else oracle.rc_pwm(mats[m]), dna, thres[m], True, mode=1)
            assert np.array_equal(h.pos32[a:b], pos.astype(np.uint32)) and np.array_equal(h.score[a:b].view(np.uint64), sc.view(np.uint64))
            assert np.all(h.motif[a:b] == m) and np.all(h.strand[a:b] == s)
    assert np.array_equal(h.pos, h.pos32.astype(np.int64)) and not h.segment.any()
    off = np.array([0, 100_000, 100_000, 350_000, 600_000], dtype=np.int64)
    hs = gpu_scan(ctx, mats, thres, dna, seg_off=off)
    assert hs.run_off is None and len(hs) > 1000
    key = (hs.segment.astype(np.int64) << 50) | (hs.motif.astype(np.int64) << 34) | (hs.strand.astype(np.int64) << 33) | hs.pos
    assert np.all(np.diff(key) > 0)


def test_ordering_dense_bucket_retry(ctx, oracle):
    """One motif hits everywhere while hundreds of others never do: the ordering pass's first bucket layout (one bucket
    per (motif, strand)) overflows the shared-memory sort and it must retry with smaller buckets — same order."""
    mats = synth.synth_pwms(300, 17, nmin=8, nmax=14)
    thres = [1e9] * len(mats)
    thres[123] = -1e9
    thres[7] = 0.5 * oracle.optimal_score(BG, mats[7])
    dna = synth.synth_dna(150_000, 3)
    h = gpu_scan(ctx, mats, thres, dna)
    n = mats[123].shape[0]
    assert len(h.select(motif=123)) == 2 * (dna.size - n + 1)
    o = oracle_scan(oracle, [mats[7], mats[123]], [thres[7], thres[123]], dna)
    assert len(h) == o[0].size
    assert np.array_equal(h.motif, np.where(o[0] == 0, 7, 123)) and np.array_equal(h.strand, o[1]) and np.array_equal(h.pos, o[2])
    assert np.array_equal(h.score.view(np.uint64), o[3].view(np.uint64))


def test_sigma_matches_oracle(ctx, oracle, ref_motifs):
    """optimalScoresSuffix (Motif/Search.hs:111-124): the library's sigma vectors (host libm) are the oracle's, bit for bit,
    for the PWM and for its reverse complement, uniform and skewed background."""
    mats = [m for _, m in ref_motifs] + synth.config_c4_motifs(40)
    for bg in (BG, (0.3, 0.2, 0.2, 0.3)):
        mo = _lib.DeviceMotifs(ctx, mats, bg)
        for m, mat in enumerate(mats):
            for s, mm in ((0, mat), (1, oracle.rc_pwm(mat))):
                assert np.array_equal(mo.get_sigma(m, s).view(np.uint64), oracle.optimal_scores_suffix(bg, mm).view(np.uint64))


def _c4_setup(ctx):
    mats = synth.config_c4_motifs(800)
    th = [float.fromhex(v) for v in json.load(open(os.path.join(GOLDEN, "config_thresholds.json")))["c4"]]
    return mats, th, _lib.Plan(_lib.DeviceMotifs(ctx, mats, BG), th)


def test_config_c4_plan_on_chunk_seam_and_n_edge(ctx, oracle):
    """BASELINE config 4 as bench.py runs it: the 800-motif plan (7 blocks, planner-chosen first-table bitmaps, 3-table
    blocks) over 8 Mbp of chr1 around the 64 Mbp item seam — scanned as the two shard items with their halo, through
    pwmscan_scan_many, merged on the host — and over 2 Mbp around the edge of the 7.5 Mbp N block; bit-identical to the
    oracle's scan of the same bytes."""
    from bioinformatics_toolkit_b200 import shard
    mats, th, plan = _c4_setup(ctx)
    assert plan.info()["blocks"] == 7
    nmax = max(m.shape[0] for m in mats)
    nlen_of = np.array([m.shape[0] for m in mats])
    L = synth.HG38_CHROM_SIZES[0][1]
    runs = [(0, 10_000), (L - 10_000, 10_000), (int(L * 0.45), int(L * 0.03))]
    # (a) the seam between the items [0, 64M) and [64M, 128M) of chr1, cut down to 4 Mbp on either side
    lo, seam, hi = 60_000_000, 64_000_000, 68_000_000
    region = synth.synth_dna(hi - lo, 4, gc=0.41, start=lo, n_runs=runs)
    items = [(0, lo, seam, 0, 800), (0, seam, hi, 0, 800)]
    arrays = []
    for it in items:
        b0, b1 = shard.chunk_bytes_range(it, hi, nmax)
        arrays.append(np.ascontiguousarray(region[b0 - lo:b1 - lo]))
    got = plan.scan_many([_lib.DeviceSeq(ctx, a) for a in arrays])
    per_item = []
    for it, h in zip(items, got):
        k = shard.keep_hit_mask(h.pos, nlen_of[h.motif], it, hi)
        per_item.append((np.zeros(int(k.sum()), np.int64), h.motif[k], h.strand[k], h.pos[k] + (it[1] - lo), h.score[k]))
    _, mo_, st_, pos_, sc_ = shard.merge_hits(per_item)
    cnt, o = oracle.scan_mt(BG, mats, th, region, mode=1, nthreads=os.cpu_count() or 8, cap=1 << 23)
    assert cnt == pos_.size and cnt > 100_000
    assert np.array_equal(mo_, o[0]) and np.array_equal(st_, o[1]) and np.array_equal(pos_, o[2])
    assert np.array_equal(sc_.view(np.uint64), o[3].view(np.uint64))
    # windows that start in the first item's last positions and end in the second item's bytes belong to the first item
    h0 = got[0]
    assert ((h0.pos < seam - lo) & (h0.pos + nlen_of[h0.motif] > seam - lo)).sum() == ((pos_ < seam - lo) & (pos_ + nlen_of[mo_] > seam - lo)).sum()
    # (b) the start of the N block (windows running into it are not hits) — through the host-bytes path
    nb = int(L * 0.45)
    lo2 = nb - 1_000_000
    reg2 = synth.synth_dna(2_000_000, 4, gc=0.41, start=lo2, n_runs=runs)
    h2 = plan.scan_host(reg2)
    cnt2, o2 = oracle.scan_mt(BG, mats, th, reg2, mode=1, nthreads=os.cpu_count() or 8, cap=1 << 22)
    assert cnt2 == len(h2) and cnt2 > 10_000
    assert np.array_equal(h2.motif, o2[0]) and np.array_equal(h2.strand, o2[1]) and np.array_equal(h2.pos, o2[2])
    assert np.array_equal(h2.score.view(np.uint64), o2[3].view(np.uint64))
    assert (h2.pos + nlen_of[h2.motif]).max() <= 1_000_000


def test_config_c2_slices_on_seams(ctx, oracle):
    """BASELINE config 2 at full size through pwmscan_scan_host: 16 oracle slices placed on the seams of the 16 Mi-base
    upload chunks and of the 32 768-position work items (plus the N-block edges) must match bit for bit."""
    dna, mats = synth.config_c2()
    th = [float.fromhex(v) for v in json.load(open(os.path.join(GOLDEN, "config_thresholds.json")))["c2"]]
    plan = _lib.Plan(_lib.DeviceMotifs(ctx, mats, BG), th)
    h = plan.scan_host(dna)
    nlen = np.array([m.shape[0] for m in mats])[h.motif]
    chunk = 16 * 1024 * 1024
    centres = [k * chunk for k in (1, 2, 3, 7, 8, 13, 14)] + [k * chunk - 32 for k in (5, 11)]            # upload-chunk seams (stored index = position + 32)
    centres += [32768 * k - 32 for k in (1, 977, 3111, 6000)] + [121_700_000, 139_700_000, dna.size - 60_000]   # work-item seams, N-block edges, the end
    assert len(centres) == 16
    tot = 0
    for c in centres:
        lo, hi = max(0, c - 60_000), min(dna.size, c + 60_000)
        cnt, o = oracle.scan_mt(BG, mats, th, dna[lo:hi], mode=1, nthreads=os.cpu_count() or 8, cap=1 << 20)
        k = (h.pos >= lo) & (h.pos + nlen <= hi)
        assert k.sum() == cnt
        assert np.array_equal(h.motif[k], o[0]) and np.array_equal(h.strand[k], o[1]) and np.array_equal(h.pos[k] - lo, o[2])
        assert np.array_equal(h.score[k].view(np.uint64), o[3].view(np.uint64))
        tot += cnt
    assert tot > 1000


def test_packed_genome_store_equals_the_bytes(ctx, oracle):
    """The packed genome store (pwmscan_pack_host -> pwmscan_seq_upload_packed / pwmscan_scan_packed_many): a sequence
    uploaded as its 2-bit + mask planes gives the hit lists of the same sequence uploaded as bytes — whole sequences
    (N runs, lower case with the upper-casing flag, lengths around the 32-base unit, empty) and slices that start at a
    multiple of 32 and end anywhere inside a longer sequence; the oracle confirms one of them."""
    mats = synth.synth_pwms(90, 8, nmin=6, nmax=21)
    thres = [0.58 * oracle.optimal_score(BG, m) for m in mats]
    plan = _lib.Plan(_lib.DeviceMotifs(ctx, mats, BG), thres)
    lens = [1_500_000, 0, 31, 32, 33, 700_001, 2_000_000]
    arrays = [synth.synth_dna(L, 500 + i, gc=0.42, n_runs=[(L // 3, 700), (L - 40, 40)] if L > 1000 else ()) for i, L in enumerate(lens)]
    low = arrays[0].copy()
    low[1000:200_000] |= 0x20                                     # lower case
    arrays.append(low)
    packed = [_lib.PackedSeq(a, uppercase=True, pinned=(i % 2 == 0)) for i, a in enumerate(arrays)]
    by_bytes = [plan.scan(_lib.DeviceSeq(ctx, a, uppercase=True)) for a in arrays]
    assert sum(len(h) for h in by_bytes) > 10000
    for a, p, ref in zip(arrays, packed, by_bytes):
        seq = _lib.DeviceSeq.from_packed(ctx, p)
        _same_hits(ref, plan.scan(seq))
        assert seq.first_invalid("Basic") == _lib.DeviceSeq(ctx, a, uppercase=True).first_invalid("Basic") == p.first_non_acgt
    many = plan.scan_packed_many([(p, 0, None) for p in packed])
    for ref, h in zip(by_bytes, many):
        _same_hits(ref, h)
    assert_same(many[5], oracle_scan(oracle, mats, thres, arrays[5]))
    # slices of a longer sequence: what the (motif block x chromosome chunk) items of a genome scan pass
    big, pbig = arrays[6], packed[6]
    cuts = [(0, 640_000 + 20), (640_000, 1_280_033), (1_280_000, 2_000_000), (1_999_968, 2_000_000), (666_656, 666_700)]
    sl = plan.scan_packed_many([(pbig, a, b) for a, b in cuts])
    for (a, b), h in zip(cuts, sl):
        _same_hits(plan.scan(_lib.DeviceSeq(ctx, big[a:b])), h)
        _same_hits(h, plan.scan(_lib.DeviceSeq.from_packed(ctx, pbig, a, b)))
    # IUPAC-aware plans need the bytes
    iupac_plan = _lib.Plan(_lib.DeviceMotifs(ctx, mats[:3], BG), thres[:3], skip_ambiguous=False)
    with pytest.raises(_lib.PwmScanError):
        iupac_plan.scan_packed_many([(packed[0], 0, None)])


def test_iupac_aware_search_through_the_first_stage(ctx, oracle):
    """skip_ambiguous = False on a sequence that is mostly A/C/G/T: the k-mer first stage finds the hits of the pure windows,
    ambig_kernel replays the windows that hold another letter (N runs, scattered IUPAC letters, the two ends of every run,
    several segments) — together exactly lookAheadSearch over every window (the oracle), and exactly the skip = True hits
    plus the ones that contain such a letter."""
    mats = synth.synth_pwms(30, 21, nmin=6, nmax=22)
    thres = [0.45 * oracle.optimal_score(BG, m) for m in mats]
    d = synth.synth_dna(300_000, 9, gc=0.45, n_runs=[(0, 70), (100_000, 3000), (299_990, 10)])
    rng = np.random.default_rng(3)
    for p in rng.integers(200, 299_000, 400):
        d[p] = rng.choice(np.frombuffer(b"NVHDBMKWSYR", dtype=np.uint8))
    h = gpu_scan(ctx, mats, thres, d, skip=False)
    o = oracle_scan(oracle, mats, thres, d, skip=False)
    assert len(h) > 5000
    assert_same(h, o)
    hs = gpu_scan(ctx, mats, thres, d, skip=True)
    assert len(hs) < len(h)
    key = lambda x: set(zip(x.motif.tolist(), x.strand.tolist(), x.pos.tolist()))
    assert key(hs) <= key(h)
    # segments: windows never span two of them, ambiguous letters right at the seams
    seg_off = [0, 50_000, 50_007, 180_000, 300_000]
    d[49_995:50_012] = ord("N")
    hseg = gpu_scan(ctx, mats, thres, d, skip=False, seg_off=seg_off)
    parts = [oracle_scan(oracle, mats, thres, d[a:b], skip=False) for a, b in zip(seg_off[:-1], seg_off[1:])]
    assert len(hseg) == sum(q[0].size for q in parts)
    for g, q in enumerate(parts):
        sel = hseg.select(segment=g)
        assert np.array_equal(sel.motif, q[0]) and np.array_equal(sel.strand, q[1]) and np.array_equal(sel.pos, q[2])
        assert np.array_equal(sel.score.view(np.uint64), q[3].view(np.uint64))


def test_max_match_sc_pruned_equals_the_maximum_of_the_dense_scores(ctx, oracle, ref_motifs):
    """maxMatchSc prunes windows against the running maximum like the reference (Search.hs:98-109) behind a first stretch of
    65 536 windows: the result is still the maximum of the exact fp64 window sums (the dense `scores` of the same sequence,
    bit for bit), for motifs of every length, with IUPAC letters and N runs in the sequence; the oracle confirms a slice."""
    mats = [m for _, m in ref_motifs][:6] + synth.synth_pwms(6, 33, nmin=5, nmax=30)
    d = synth.synth_dna(3_000_000, 21, gc=0.43, n_runs=[(1_000_000, 5000)])
    rng = np.random.default_rng(5)
    for p in rng.integers(0, d.size, 300):
        d[p] = rng.choice(np.frombuffer(b"NVHDBMKWSYR", dtype=np.uint8))
    seq = _lib.DeviceSeq(ctx, d, keep_ascii=True)
    mo = _lib.DeviceMotifs(ctx, mats, BG)
    mx = _lib.max_match_sc(seq, mo)
    for m in range(len(mats)):
        dense = _lib.scores(seq, mo, m, 0)
        assert mx[m] == dense.max() and np.float64(mx[m]).view(np.uint64) == dense.max().view(np.uint64)
    sl = bytes(d[900_000:1_200_000])
    mxs = _lib.max_match_sc(_lib.DeviceSeq(ctx, sl, keep_ascii=True), mo)
    for m in (0, 5, len(mats) - 1):
        assert mxs[m] == oracle.max_match_sc(BG, mats[m], sl)
    bad = d.copy()
    bad[2_999_990] = ord("!")                      # an invalid letter is still the reference's `error`
    with pytest.raises(_lib.InvalidNucleotide):
        _lib.max_match_sc(_lib.DeviceSeq(ctx, bad, keep_ascii=True), mo)
