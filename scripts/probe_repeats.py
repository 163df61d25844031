"""How does the scan behave on low-complexity sequence?  C2 motifs x 64 Mbp, with 0 / 5 / 20 % of the sequence
overwritten by tandem repeats ((A)n, (CA)n, (GGAA)n, (TTAGGG)n, period-171 satellite copies)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bioinformatics_toolkit_b200 import _lib, synth
BG = (0.25,) * 4
ctx = _lib.default_context(0)
mats = synth.synth_pwms(100, 12)
th = np.array([float.fromhex(v) for v in json.load(open("tests/golden/config_thresholds.json"))["c2"]])
mo = _lib.DeviceMotifs(ctx, mats, BG)
plan = _lib.Plan(mo, th)
L = 64_000_000
base = synth.synth_dna(L, 5, gc=0.41)
rng = np.random.default_rng(3)
units = [b"A", b"CA", b"GGAA", b"TTAGGG", bytes(synth.synth_dna(171, 9))]
for frac in (0.0, 0.05, 0.20):
    dna = base.copy()
    covered = 0
    while covered < frac * L:
        u = np.frombuffer(units[int(rng.integers(0, len(units)))], dtype=np.uint8)
        n = int(rng.integers(200, 20000))
        s = int(rng.integers(0, L - n))
        dna[s:s + n] = np.resize(u, n)
        covered += n
    seq = _lib.DeviceSeq(ctx, dna)
    for rep in range(3):
        t0 = time.perf_counter(); h = plan.scan(seq); t1 = time.perf_counter()
    tm = ctx.timing()
    print("repeats %4.0f %%: scan %.2f ms, prefilter %.2f ms, candidates %d, hits %d" % (100 * frac, 1e3 * (t1 - t0), tm["prefilter_ms"], tm["candidates"], len(h)))
    del h, seq
