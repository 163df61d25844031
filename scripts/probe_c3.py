"""Where does a scanMotif-sized call spend its time?  Plan creation and one pipelined scan of
50 000 x 500-bp segments x 800 motifs, per first-stage mode (PWMSCAN_PREFILTER)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bioinformatics_toolkit_b200 import _lib, synth
BG = (0.25,) * 4
ctx = _lib.default_context(0)
mats = synth.config_c4_motifs(800)
th = np.array([float.fromhex(v) for v in json.load(open("tests/golden/config_thresholds.json"))["c4"]])
dna = synth.synth_dna(25_000_000, 5, gc=0.41)
off = np.arange(0, 25_000_001, 500, dtype=np.int64)
t0 = time.perf_counter(); mo = _lib.DeviceMotifs(ctx, mats, BG); t1 = time.perf_counter()
print("motifs_create %.1f ms" % (1e3 * (t1 - t0)))
for rep in range(3):
    t0 = time.perf_counter(); plan = _lib.Plan(mo, th, strands=2, skip_ambiguous=True); t1 = time.perf_counter()
    h = plan.scan_host(dna, seg_off=off); t2 = time.perf_counter()
    tm = ctx.timing()
    h2 = plan.scan_host(dna, seg_off=off); t3 = time.perf_counter()
    print("plan %.1f ms | first scan %.1f ms (device %s) | second scan %.1f ms | hits %d" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), {k: round(v, 2) for k, v in tm.items() if k.endswith("_ms")}, 1e3 * (t3 - t2), len(h)))
    t4 = time.perf_counter(); h3 = plan.scan_host(dna, seg_off=off, uppercase=True); t5 = time.perf_counter()
    del h, h2, h3
    t6 = time.perf_counter(); del plan; t7 = time.perf_counter()
    print("   uppercase scan %.1f ms | plan free %.1f ms" % (1e3 * (t5 - t4), 1e3 * (t7 - t6)))
