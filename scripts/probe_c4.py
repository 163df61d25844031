"""One C4-sized chunk (64 Mbp x 800 motifs, both strands): device timing breakdown of a resident scan."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bioinformatics_toolkit_b200 import _lib, synth
BG = (0.25,) * 4
ctx = _lib.default_context(0)
mats = synth.config_c4_motifs(800)
th = np.array([float.fromhex(v) for v in json.load(open("tests/golden/config_thresholds.json"))["c4"]])
dna = synth.synth_dna(int(os.environ.get("PROBE_BP", "64000000")), 5, gc=0.41)
mo = _lib.DeviceMotifs(ctx, mats, BG)
plan = _lib.Plan(mo, th)
seq = _lib.DeviceSeq(ctx, dna)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    t0 = time.perf_counter(); h = plan.scan(seq); t1 = time.perf_counter()
    tm = ctx.timing()
    print("scan %.2f ms | %s | hits %d cand %d kernels %d d2h %.1f MB" % (1e3 * (t1 - t0), {k: round(float(v), 3) for k, v in tm.items() if k.endswith("_ms")}, len(h),
          tm["candidates"], tm["kernel_launches"], tm["d2h_bytes"] / 1e6))
    del h
