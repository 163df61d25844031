"""One 128-motif block of the config-4 set (motifs sorted by length, block index argv[1]) on a 64 Mbp chunk: for ncu."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bioinformatics_toolkit_b200 import _lib, synth
BG = (0.25,) * 4
ctx = _lib.default_context(0)
mats = synth.config_c4_motifs(800)
th = np.array([float.fromhex(v) for v in json.load(open("tests/golden/config_thresholds.json"))["c4"]])
b = int(sys.argv[1]) if len(sys.argv) > 1 else 1
order = sorted(range(800), key=lambda i: mats[i].shape[0])
idx = order[128 * b: 128 * b + 128]
seq = _lib.DeviceSeq(ctx, synth.synth_dna(64_000_000, 5, gc=0.41))
plan = _lib.Plan(_lib.DeviceMotifs(ctx, [mats[i] for i in idx], BG), th[idx])
for rep in range(3):
    h = plan.scan(seq); print(len(h), ctx.timing()["prefilter_ms"]); del h
