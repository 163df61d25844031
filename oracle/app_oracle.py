"""Python restatements (TEST INFRASTRUCTURE ONLY — same status as oracle/pwm_oracle.cpp) of the
host-side pieces of the reference that sit around the C++ oracle kernels:

  scan_motif_lines   MotifScan.hs:36-42 + Data/Bed/Utils.hs:88-124 (mkCutoffMotif, scanMotif, toLine)
  iterative_merge    Motif/Merge.hs:113-170 (with the oracle's alignment), mergePWMWeighted 48-71

Plain loops, small inputs only.  PARITY STATUS: unpinned (the reference has no test for these).
"""
import numpy as np

from . import oracle as O


def mk_cutoff_motif(bg, p, mat):
    cdf = O.score_cdf(bg, mat)
    return {"mat": mat, "rc": O.rc_pwm(mat), "cutoff": O.cdf_inv(cdf, 1 - p), "cdf": O.truncate_cdf(1 - p * 10, cdf),
            "sigma": O.optimal_scores_suffix(bg, mat), "sigma_rc": O.optimal_scores_suffix(bg, O.rc_pwm(mat))}


def scan_motif_lines(genome, names, mats, beds, p=1e-5, bg=O.DEF_BG):
    """genome: {chrom(bytes): uint8 array}.  Returns the BED6 lines scan_motif prints."""
    cms = [mk_cutoff_motif(bg, p, m) for m in mats]
    iupac = set(b"ACGTNVHDBMKWSYR")
    lines = []
    for (chr_, start, end) in beds:
        if chr_ not in genome or end > len(genome[chr_]):
            continue                                              # Left _ -> warning, region skipped
        dna = bytes(genome[chr_][start:end]).upper()
        if any(c not in iupac for c in dna):
            continue
        for name, cm in zip(names, cms):
            n = cm["mat"].shape[0]
            for strand, (mat, sg) in ((b"+", (cm["mat"], cm["sigma"])), (b"-", (cm["rc"], cm["sigma_rc"]))):
                pos, sc = O.find_tfbs(bg, mat, dna, cm["cutoff"], True, sigma=sg)
                for i, s in zip(pos, sc):
                    aff = O.to_affinity(1 - O.cdf(cm["cdf"], s))
                    lines.append(b"\t".join((chr_, b"%d" % (start + i), b"%d" % (start + i + n), name, b"%d" % aff, strand)))
    return lines


def scan_motif_lines_mt(genome, names, mats, beds, p=1e-5, bg=O.DEF_BG, nthreads=8):
    """Same lines as scan_motif_lines for many regions x many motifs: per region ONE multi-threaded oracle scan of all
    motifs and both strands (O.scan_mt, hits in (motif, strand, position) order = scanMotif's order within a region)
    instead of one call per (region, motif, strand).  Checked against scan_motif_lines in tests/test_oracle_invariants.py."""
    cms = [mk_cutoff_motif(bg, p, m) for m in mats]
    th = [cm["cutoff"] for cm in cms]
    nlen = [cm["mat"].shape[0] for cm in cms]
    iupac = set(b"ACGTNVHDBMKWSYR")
    lines = []
    for (chr_, start, end) in beds:
        if chr_ not in genome or end > len(genome[chr_]):
            continue
        dna = bytes(genome[chr_][start:end]).upper()
        if any(c not in iupac for c in dna):
            continue
        cnt, o = O.scan_mt(bg, mats, th, np.frombuffer(dna, dtype=np.uint8), mode=1, nthreads=nthreads, cap=1 << 16)
        for m, s, i, sc in zip(o[0], o[1], o[2], o[3]):
            aff = O.to_affinity(1 - O.cdf(cms[m]["cdf"], sc))
            lines.append(b"\t".join((chr_, b"%d" % (start + i), b"%d" % (start + i + nlen[m]), names[m], b"%d" % aff, b"-" if s else b"+")))
    return lines


def merge_pwm_weighted(m1, m2, shift):
    (p1, w1), (p2, w2), s = (m1, m2, shift) if shift >= 0 else (m2, m1, -shift)
    n1, n2 = p1.shape[0], p2.shape[0]
    rows, ws, i = [], [], 0
    while True:
        i2 = i - s
        if i2 < 0 or (i < n1 and i2 >= n2):
            rows.append(p1[i]); ws.append(w1[i])
        elif i < n1 and i2 < n2:
            rows.append((float(w1[i]) * p1[i] + float(w2[i2]) * p2[i2]) / float(w1[i] + w2[i2])); ws.append(w1[i] + w2[i2])
        elif i >= n1 and i2 < n2:
            rows.append(p2[i2]); ws.append(w2[i2])
        else:
            break
        i += 1
    return np.array(rows), ws


def iterative_merge(th, names, mats, penal="exp", gap=0.05, comb="l1"):
    n = len(mats)
    items = [([nm], m, [1] * m.shape[0]) for nm, m in zip(names, mats)]
    mat = {}
    for i in range(n):
        for j in range(i + 1, n):
            mat[(i, j)] = O.alignment(items[i][1], items[j][1], penal, gap, comb)
    while True:
        best, bd = None, (float("inf"), None)
        for i in range(n):
            for j in range(i + 1, n):
                x = mat.get((i, j))
                if x is not None and x[0] < bd[0]:
                    best, bd = (i, j), x
        if best is None or not (bd[0] < th):
            break
        (i, j), (d, (same, pos)) = best, bd
        nm1, p1, w1 = items[i]
        nm2, p2, w2 = items[j]
        pw, ww = merge_pwm_weighted((p1, w1), (p2, w2), pos) if same else merge_pwm_weighted((p1, w1), (O.rc_pwm(p2), w2[::-1]), pos)
        for q in range(n):
            mat[(min(q, j), max(q, j))] = None
        items[i] = (nm1 + nm2, pw, ww)
        items[j] = None
        for q in range(n):
            if q != i and items[q] is not None:
                mat[(min(i, q), max(i, q))] = O.alignment(pw, items[q][1], penal, gap, comb) if i < q else O.alignment(items[q][1], pw, penal, gap, comb)
    return [it for it in items if it is not None]
