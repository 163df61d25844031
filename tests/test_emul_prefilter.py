"""CPU check of the device pipeline's host-visible logic: the planning code (csrc/plan.h) and the
per-lane kernel body (csrc/scan_core.cuh) are compiled with g++ and run lane by lane
(tests/emul/emul_scan.cpp); the hit list must equal the oracle's bit for bit — in particular the
quantised prefilter must never lose a hit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from bioinformatics_toolkit_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
BG = (0.25, 0.25, 0.25, 0.25)


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(HERE, "emul", "libemul.so")
    src = os.path.join(HERE, "emul", "emul_scan.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-w", "-shared", "-o", so, src])
    L = C.CDLL(so)
    L.emul_scan.restype = C.c_int64

    def run(mats, thres, dna, strands=2, force_s1=0, bg=BG):
        ncol = np.array([m.shape[0] for m in mats], dtype=np.int32)
        flat = np.concatenate([np.ascontiguousarray(m, dtype=np.float64).reshape(-1) for m in mats])
        cap = 1 << 21
        om, os_, op, osc = np.empty(cap, np.int32), np.empty(cap, np.uint8), np.empty(cap, np.int64), np.empty(cap)
        st = np.zeros(8, np.int64)
        bga, th = np.array(bg, float), np.array(thres, float)
        d = np.ascontiguousarray(np.frombuffer(bytes(dna), dtype=np.uint8))
        vp = C.c_void_p
        n = L.emul_scan(C.c_int(len(mats)), ncol.ctypes.data_as(vp), flat.ctypes.data_as(vp), bga.ctypes.data_as(vp),
                        th.ctypes.data_as(vp), C.c_int(strands), d.ctypes.data_as(vp), C.c_int64(d.size), C.c_int(force_s1),
                        om.ctypes.data_as(vp), os_.ctypes.data_as(vp), op.ctypes.data_as(vp), osc.ctypes.data_as(vp),
                        C.c_int64(cap), st.ctypes.data_as(vp))
        assert 0 <= n <= cap
        return (om[:n].copy(), os_[:n].copy(), op[:n].copy(), osc[:n].copy()), st
    return run


def _same(h, o):
    return all(np.array_equal(a, b) for a, b in zip(h, o))


@pytest.mark.parametrize("force_s1", [0, 2, 3, 4])
def test_mixed_lengths_both_strands(emul, oracle, force_s1):
    mats = synth.synth_pwms(70, 21, nmin=6, nmax=24)
    dna = synth.synth_dna(60000, 9, gc=0.45, n_runs=[(0, 40), (30000, 700), (59990, 10)])
    thres = [0.55 * oracle.optimal_score(BG, m) for m in mats]
    h, st = emul(mats, thres, dna, force_s1=force_s1)
    cnt, o = oracle.scan_mt(BG, mats, thres, dna, mode=1, nthreads=4, cap=1 << 21)
    assert cnt == len(o[0]) and cnt > 100
    assert _same(h, o)
    assert st[2] == 2 and st[3] == 0


def test_pvalue_thresholds_and_reference_fixtures(emul, oracle, ref_motifs, ref_motifs_meme):
    mats = [m for _, m in ref_motifs] + [m for _, m in ref_motifs_meme]
    dna = synth.synth_dna(200000, 3)
    thres = [oracle.pvalue_to_score(1e-4, BG, m) for m in mats]
    h, st = emul(mats, thres, dna)
    cnt, o = oracle.scan_mt(BG, mats, thres, dna, mode=0, nthreads=4, cap=1 << 20)
    assert cnt > 50 and _same(h, o)
    assert st[0] < 4 * cnt + 100          # the prefilter is tight: few false candidates


def test_long_motifs_go_to_the_exact_path(emul, oracle):
    mats = synth.synth_pwms(3, 5, nmin=29, nmax=40) + synth.synth_pwms(3, 6, nmin=25, nmax=28)
    dna = synth.synth_dna(20000, 4)
    thres = [0.35 * oracle.optimal_score(BG, m) for m in mats]
    h, st = emul(mats, thres, dna)
    cnt, o = oracle.scan_mt(BG, mats, thres, dna, mode=1, nthreads=2, cap=1 << 20)
    assert st[3] == 6 and _same(h, o)


def test_forward_only_nonuniform_background_and_edge_cutoffs(emul, oracle):
    bg = (0.3, 0.2, 0.2, 0.3)
    mats = synth.synth_pwms(9, 8, nmin=5, nmax=16)
    dna = synth.synth_dna(50000, 6, gc=0.4)
    opt = [max(oracle.optimal_score(bg, m), 1.0) for m in mats]
    thres = [0.5 * opt[0], 0.9 * opt[1], 1.5 * opt[2], -1e9, 0.0, 0.7 * opt[5], float("inf"), 0.6 * opt[7], 0.3 * opt[8]]
    mats[3] = mats[3][:5]
    h, st = emul(mats, thres, dna[:3000], strands=1, bg=bg)
    ref = []
    for m, (mat, th) in enumerate(zip(mats, thres)):
        pos, sc = oracle.find_tfbs(bg, mat, dna[:3000], th, True)
        ref.append((np.full(pos.size, m, np.int32), np.zeros(pos.size, np.uint8), pos, sc))
    o = tuple(np.concatenate([r[k] for r in ref]) for k in range(4))
    assert _same(h, o)
    assert (h[0] == 3).sum() == 3000 - 5 + 1          # cutoff -1e9: every window is a hit
    assert (h[0] == 2).sum() == 0 and (h[0] == 6).sum() == 0


def test_tiny_and_empty_sequences(emul, oracle, ref_motifs):
    mats = [m for _, m in ref_motifs]
    thres = [0.0] * len(mats)
    for dna in (b"", b"ACGT", b"ACGTACGTAC", bytes(synth.synth_dna(11, 1)), bytes(synth.synth_dna(33, 1))):
        h, _ = emul(mats, thres, dna)
        cnt, o = oracle.scan_mt(BG, mats, thres, np.frombuffer(dna, dtype=np.uint8), mode=1, cap=1 << 16)
        if cnt == 0:
            assert h[0].size == 0
        else:
            assert _same(h, o)


@pytest.fixture(scope="module")
def emul_mma(emul):
    L = C.CDLL(os.path.join(HERE, "emul", "libemul.so"))
    L.emul_mma_scan.restype = C.c_int64

    def run(mats, thres, dna, strands=2, force_steps=0, bg=BG):
        ncol = np.array([m.shape[0] for m in mats], dtype=np.int32)
        flat = np.concatenate([np.ascontiguousarray(m, dtype=np.float64).reshape(-1) for m in mats])
        cap = 1 << 21
        om, os_, op, osc = np.empty(cap, np.int32), np.empty(cap, np.uint8), np.empty(cap, np.int64), np.empty(cap)
        st = np.zeros(8, np.int64)
        bga, th = np.array(bg, float), np.array(thres, float)
        d = np.ascontiguousarray(np.frombuffer(bytes(dna), dtype=np.uint8))
        vp = C.c_void_p
        n = L.emul_mma_scan(C.c_int(len(mats)), ncol.ctypes.data_as(vp), flat.ctypes.data_as(vp), bga.ctypes.data_as(vp),
                            th.ctypes.data_as(vp), C.c_int(strands), d.ctypes.data_as(vp), C.c_int64(d.size), C.c_int(force_steps),
                            om.ctypes.data_as(vp), os_.ctypes.data_as(vp), op.ctypes.data_as(vp), osc.ctypes.data_as(vp),
                            C.c_int64(cap), st.ctypes.data_as(vp))
        assert 0 <= n <= cap
        return (om[:n].copy(), os_[:n].copy(), op[:n].copy(), osc[:n].copy()), st
    return run


@pytest.mark.parametrize("force_steps", [0, 1, 2, 4])
def test_mma_plan_never_loses_a_hit(emul_mma, oracle, force_steps):
    """csrc/plan_mma.h: the int8 weights of the tensor-core first stage, summed with plain integer
    loops over the same layout the MMA reads, must flag every window the oracle reports."""
    mats = synth.synth_pwms(40, 23, nmin=5, nmax=40)
    dna = synth.synth_dna(30000, 10, gc=0.45, n_runs=[(0, 30), (12000, 500), (29990, 10)])
    thres = [0.55 * oracle.optimal_score(BG, m) for m in mats]
    thres[3] = -1e9
    thres[4] = 2.0 * abs(thres[4]) + 50
    h, st = emul_mma(mats, thres, dna[:6000], force_steps=force_steps)
    cnt, o = oracle.scan_mt(BG, mats, thres, dna[:6000], mode=1, nthreads=4, cap=1 << 21)
    assert cnt > 100 and _same(h, o)
    assert st[3] == 0          # no motif-length limit on this path


def test_mma_plan_pvalue_cutoffs_are_tight(emul_mma, oracle, ref_motifs, ref_motifs_meme):
    mats = [m for _, m in ref_motifs] + [m for _, m in ref_motifs_meme]
    dna = synth.synth_dna(60000, 3)
    thres = [oracle.pvalue_to_score(1e-4, BG, m) for m in mats]
    h, st = emul_mma(mats, thres, dna, strands=1)
    ref = []
    for m, (mat, th) in enumerate(zip(mats, thres)):
        pos, sc = oracle.find_tfbs(BG, mat, dna, th, True)
        ref.append((np.full(pos.size, m, np.int32), np.zeros(pos.size, np.uint8), pos, sc))
    o = tuple(np.concatenate([r[k] for r in ref]) for k in range(4))
    assert _same(h, o) and len(o[0]) > 10
    assert st[0] < 30 * len(o[0]) + 200


@pytest.fixture(scope="module")
def emul_km(emul):
    L = C.CDLL(os.path.join(HERE, "emul", "libemul.so"))
    L.emul_km_scan.restype = C.c_int64

    def run(mats, thres, dna, strands=2, force_nt=0, force_L=0, bg=BG):
        ncol = np.array([m.shape[0] for m in mats], dtype=np.int32)
        flat = np.concatenate([np.ascontiguousarray(m, dtype=np.float64).reshape(-1) for m in mats])
        cap = 1 << 21
        om, os_, op, osc = np.empty(cap, np.int32), np.empty(cap, np.uint8), np.empty(cap, np.int64), np.empty(cap)
        st = np.zeros(8, np.int64)
        bga, th = np.array(bg, float), np.array(thres, float)
        d = np.ascontiguousarray(np.frombuffer(bytes(dna), dtype=np.uint8))
        vp = C.c_void_p
        n = L.emul_km_scan(C.c_int(len(mats)), ncol.ctypes.data_as(vp), flat.ctypes.data_as(vp), bga.ctypes.data_as(vp),
                           th.ctypes.data_as(vp), C.c_int(strands), d.ctypes.data_as(vp), C.c_int64(d.size), C.c_int(force_nt), C.c_int(force_L),
                           om.ctypes.data_as(vp), os_.ctypes.data_as(vp), op.ctypes.data_as(vp), osc.ctypes.data_as(vp),
                           C.c_int64(cap), st.ctypes.data_as(vp))
        assert 0 <= n <= cap
        return (om[:n].copy(), os_[:n].copy(), op[:n].copy(), osc[:n].copy()), st
    return run


@pytest.mark.parametrize("force_nt,force_L", [(0, 0), (1, 8), (2, 9), (3, 10), (4, 8)])
def test_km_plan_never_loses_a_hit(emul_km, oracle, force_nt, force_L):
    """csrc/plan_km.h: the k-mer mask bits (window deficit <= Dsafe, the sum km_build_kernel forms) and the
    12-bit middle stage, evaluated with plain loops, must keep every window the oracle reports — for every
    table layout, motifs of 5..40 columns, N runs, an unreachable and an always-true cutoff."""
    mats = synth.synth_pwms(40, 23, nmin=5, nmax=40)
    dna = synth.synth_dna(30000, 10, gc=0.45, n_runs=[(0, 30), (12000, 500), (29990, 10)])
    thres = [0.55 * oracle.optimal_score(BG, m) for m in mats]
    thres[3] = -1e9
    thres[4] = 2.0 * abs(thres[4]) + 50
    h, st = emul_km(mats, thres, dna[:5000], force_nt=force_nt, force_L=force_L)
    cnt, o = oracle.scan_mt(BG, mats, thres, dna[:5000], mode=1, nthreads=4, cap=1 << 21)
    assert cnt > 100 and _same(h, o)
    assert st[3] == 0                                    # no motif-length limit on this path
    if force_nt:
        assert st[4] == force_nt and st[5] == force_L


def test_km_plan_pvalue_cutoffs_are_selective(emul_km, oracle, ref_motifs, ref_motifs_meme):
    """At p-value cutoffs the two stages must be tight: candidates close to the hit count, first-stage
    survivors within a small factor of the planner's own estimate."""
    mats = [m for _, m in ref_motifs] + [m for _, m in ref_motifs_meme]
    dna = synth.synth_dna(60000, 3)
    thres = [oracle.pvalue_to_score(1e-4, BG, m) for m in mats]
    h, st = emul_km(mats, thres, dna, strands=1)
    ref = []
    for m, (mat, th) in enumerate(zip(mats, thres)):
        pos, sc = oracle.find_tfbs(BG, mat, dna, th, True)
        ref.append((np.full(pos.size, m, np.int32), np.zeros(pos.size, np.uint8), pos, sc))
    o = tuple(np.concatenate([r[k] for r in ref]) for k in range(4))
    assert _same(h, o) and len(o[0]) > 10
    assert st[0] < 1.2 * len(o[0]) + 20                  # middle stage: candidates ~ hits
    assert st[6] < 3.0 * st[7] * 1e-6 * dna.size + 200   # first stage: survivors ~ the plan's estimate


def test_km_planner_decisions(emul):
    """The k-mer mask planner on the benchmark motif sets (thresholds = the committed p < 1e-5 fixtures): C2's single
    dense block gathers two 10-mer tables without the bitmap; the many-motif set (sorted by length) gets sparse blocks
    with the first-table bitmap and far fewer sectors than one per block and position; every layout respects the
    kernel's limits (k-mers of all tables within 32 bases, tables <= 72 MiB per block, shared-memory budget)."""
    import json
    L = C.CDLL(os.path.join(HERE, "emul", "libemul.so"))
    th_all = json.load(open(os.path.join(HERE, "golden", "config_thresholds.json")))

    def plan(mats, thres, strands=2):
        ncol = np.array([m.shape[0] for m in mats], dtype=np.int32)
        flat = np.concatenate([np.ascontiguousarray(m, dtype=np.float64).reshape(-1) for m in mats])
        out = np.zeros(8 * 64, np.int64)
        vp = C.c_void_p
        nb = L.emul_km_plan_info(C.c_int(len(mats)), ncol.ctypes.data_as(vp), flat.ctypes.data_as(vp), np.array(BG, float).ctypes.data_as(vp),
                                 np.array(thres, float).ctypes.data_as(vp), C.c_int(strands), out.ctypes.data_as(vp), C.c_int(64))
        assert nb >= 0, "a block violates the kernel's layout limits"
        return out[:8 * nb].reshape(nb, 8)

    c2 = plan(synth.synth_pwms(100, 12), [float.fromhex(v) for v in th_all["c2"]])
    assert c2.shape[0] == 1
    nt, l0, bitmap, sectors, surv, cand, max_n, mib = c2[0]
    assert (nt, l0, bitmap) == (2, 10, 0) and 1.5e6 < sectors < 1.8e6 and surv < 0.06e6 and 1.5e3 < cand < 3e3 and mib == 64
    c4 = plan(synth.config_c4_motifs(800), [float.fromhex(v) for v in th_all["c4"]])
    assert c4.shape[0] == 7                                    # 800 motifs / 128 per block
    assert c4[:, 2].sum() >= 5                                 # sparse first tables -> bitmap (the block of the longest motifs is dense: plain path)
    assert c4[:, 3].sum() < 2.0e6                              # < 2 sectors per position over all seven blocks
    assert np.all(np.diff(c4[:, 6]) <= 0)                      # blocks are sorted by motif length, longest first
    one = plan([synth.ctcf_like_pwm()], [float.fromhex(th_all["c1"][0])])
    assert one.shape[0] == 1 and one[0, 3] < 0.05e6            # one motif: almost every position stops at the bitmap


def test_km_plan_property_random_cases(emul_km, oracle):
    """Randomised sweep (fixed seeds): motif counts, lengths 1..64, skewed backgrounds, one or both strands, cutoffs
    from very loose to consensus-only, forced and free table layouts — the k-mer mask plan + middle stage never lose
    or invent a hit relative to the oracle."""
    rng = np.random.default_rng(2024)
    layouts = [(0, 0), (1, 8), (1, 10), (2, 8), (2, 10), (3, 9), (3, 10), (4, 8)]
    for case in range(24):
        nm = int(rng.integers(1, 9))
        lengths = [int(rng.choice([1, 2, 3, 5, 8, 9, 10, 11, 16, 17, 21, 31, 33, 48, 64])) for _ in range(nm)]
        mats = synth.synth_pwms(nm, 500 + case, lengths=lengths, alpha=float(rng.choice([0.15, 0.3, 1.0])), zero_frac=float(rng.choice([0.0, 0.125, 0.4])))
        bgw = rng.dirichlet([8, 8, 8, 8])
        bg = tuple(float(x) for x in bgw / bgw.sum())
        dna = synth.synth_dna(int(rng.integers(50, 2500)), 900 + case, gc=float(rng.uniform(0.3, 0.7)),
                              n_runs=[(int(rng.integers(0, 40)), int(rng.integers(1, 30)))] if rng.random() < 0.5 else [])
        thres = []
        for m in mats:
            opt = oracle.optimal_score(bg, m)
            thres.append(float(rng.choice([-1e9, -2.0 * abs(opt), -0.5 * abs(opt), 0.0, 0.3 * opt, 0.7 * opt, opt - 1e-9, opt, opt + 1.0])))
        strands = int(rng.integers(1, 3))
        nt, L = layouts[case % len(layouts)]
        h, st = emul_km(mats, thres, dna, strands=strands, force_nt=nt, force_L=L, bg=bg)
        ref = []
        for mi, (mat, th) in enumerate(zip(mats, thres)):
            for s in range(strands):
                mm = mat if s == 0 else oracle.rc_pwm(mat)
                pos, sc = oracle.find_tfbs(bg, mm, dna, th, True, mode=1)
                ref.append((np.full(pos.size, mi, np.int32), np.full(pos.size, s, np.uint8), pos, sc))
        o = tuple(np.concatenate([r[k] for r in ref]) for k in range(4))
        assert _same(h, o), (case, lengths, thres, strands, nt, L)
