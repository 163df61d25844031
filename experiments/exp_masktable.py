"""Selectivity of 1-bit 'alive mask' L-mer tables (MOODS-style lookahead filtration) on the C2 / C4 motif sets.
For each combo: windows of L columns; sigma_w = P(deficit in window <= Dsafe) under iid uniform bases.
Reports sum over combos of: best single window, product over a partition into consecutive windows."""
import sys, json, numpy as np
sys.path.insert(0, '.')
from oracle import oracle as O
from bioinformatics_toolkit_b200 import synth
bg = O.DEF_BG
which = sys.argv[1] if len(sys.argv) > 1 else 'c2'
mats = synth.synth_pwms(100, 12) if which == 'c2' else synth.config_c4_motifs(800)[:200]
thres = [float.fromhex(v) for v in json.load(open('tests/golden/config_thresholds.json'))[which][:len(mats)]]

def window_defs(d):            # d: L x 4 deficits -> all 4^L sums
    s = np.zeros(1)
    for row in d:
        s = (s[:, None] + row[None, :]).reshape(-1)
    return s

for L in (6, 8, 10):
    tot_single = 0.0; tot_and = 0.0; tot_stair = 0.0; nact = 0
    per = []
    for m, th in zip(mats, thres):
        n = m.shape[0]
        for strand in (0, 1):
            mm = m if strand == 0 else O.rc_pwm(m)
            S = O.score_table(bg, mm)
            D = S.max(axis=1).sum() - th
            if not np.isfinite(th) or D <= 0:
                continue
            nact += 1
            d = S.max(axis=1, keepdims=True) - S
            # best single window
            best = 1.0
            for c in range(0, max(1, n - L + 1)):
                w = window_defs(d[c:c + L])
                best = min(best, (w <= D).mean())
            tot_single += best
            # partition from offset c0 in consecutive windows (ragged ends allowed), product of per-window pass rates
            bestp = 1.0; bests = 1.0
            for c0 in range(0, L):
                cuts = [0] if c0 == 0 else [0, c0]
                while cuts[-1] + L < n: cuts.append(cuts[-1] + L)
                cuts.append(n)
                ws = [window_defs(d[a:b]) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
                p = 1.0
                for w in ws: p *= (w <= D).mean()
                bestp = min(bestp, p)
            tot_and += bestp
            per.append((n, best, bestp))
    per = np.array(per)
    print(f'L={L}: combos {nact}  sum sigma single-window {tot_single:.4f}  AND-of-windows {tot_and:.5f}  (true hits/pos ~ {nact*1e-5:.5f})')
    for lo, hi in ((6, 9), (10, 12), (13, 16), (17, 22)):
        sel = (per[:, 0] >= lo) & (per[:, 0] <= hi)
        if sel.any(): print(f'   n {lo}-{hi}: {sel.sum()} combos, single {per[sel,1].mean():.2e}, and {per[sel,2].mean():.2e}')
