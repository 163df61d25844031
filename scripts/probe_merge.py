"""iterativeMerge (Merge.hs:113-170) on the C5 motif set: all-pairs matrix on the device, greedy merge loop."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bioinformatics_toolkit_b200 import motif as Mo, merge as M, synth
from bioinformatics_toolkit_b200.alignment import alignment
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
mats = synth.config_c5_motifs(n)
motifs = [Mo.Motif(b"m%d" % i, Mo.PWM(20, m)) for i, m in enumerate(mats)]
for th in (0.05, 0.2):
    t0 = time.perf_counter()
    out = M.iterativeMerge(alignment, th, motifs)
    print("iterativeMerge n=%d th=%.2f: %.2f s -> %d motifs" % (n, th, time.perf_counter() - t0, len(out)))
