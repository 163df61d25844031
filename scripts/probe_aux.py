"""Timing of the entry points next to the main path: IUPAC-aware scan (skip_ambiguous = False: exact_kernel over every window),
dense scores / maxMatchSc (scores_kernel, max_kernel), per-hit affinity (affinity_kernel)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bioinformatics_toolkit_b200 import _lib, synth
from bioinformatics_toolkit_b200 import motif as Mo
BG = (0.25,) * 4
ctx = _lib.default_context(0)
mats = synth.synth_pwms(100, 12)
th = np.array([float.fromhex(v) for v in json.load(open("tests/golden/config_thresholds.json"))["c2"]])
L = 16_000_000
dna = synth.synth_dna(L, 7, gc=0.41, n_runs=[(5_000_000, 20000)])
dna[1000:1010] = np.frombuffer(b"RYKMSWBDHV", dtype=np.uint8)         # IUPAC letters the aware search has to score
mo = _lib.DeviceMotifs(ctx, mats, BG)
seq = _lib.DeviceSeq(ctx, dna, keep_ascii=True)
def timed(label, fn, reps=3, unit=None):
    fn(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
    tm = ctx.timing()
    print("%-58s %8.2f ms (device: %s)%s" % (label, 1e3 * min(ts), {k: round(float(v), 3) for k, v in tm.items() if k.endswith("_ms") and v}, unit(min(ts)) if unit else ""), flush=True)
    return r
plan_skip = _lib.Plan(mo, th)
plan_iupac = _lib.Plan(mo, th, skip_ambiguous=False)
h1 = timed("findTFBS skip=True, 100 PWMs x 16 Mbp", lambda: plan_skip.scan(seq), unit=lambda t: "  %.0f Gbp*motifs/s" % (L * 100 / t / 1e9))
h2 = timed("findTFBS skip=False (IUPAC-aware), 100 PWMs x 16 Mbp", lambda: plan_iupac.scan(seq), unit=lambda t: "  %.0f Gbp*motifs/s" % (L * 100 / t / 1e9))
print("   hits skip=True %d, skip=False %d" % (len(h1), len(h2)))
timed("scores' of one 20-column PWM over 16 Mbp (8 B per window)", lambda: _lib.scores(seq, mo, 3), unit=lambda t: "  %.1f G windows/s" % (L / t / 1e9))
timed("maxMatchSc of 100 PWMs over 16 Mbp", lambda: _lib.max_match_sc(seq, mo), unit=lambda t: "  %.0f Gbp*motifs/s" % (L * 100 / t / 1e9))
_, cdfs = mo.cutoff_cdfs(1e-5)
off = np.concatenate([[0], np.cumsum([c[0].size for c in cdfs])]).astype(np.int64)
xs = np.concatenate([c[0] for c in cdfs]); ys = np.concatenate([c[1] for c in cdfs])
timed("per-hit p-value + toAffinity of %d hits" % len(h1), lambda: h1.affinity(off, xs, ys), unit=lambda t: "  %.1f M hits/s" % (len(h1) / t / 1e6))
