"""Score CDFs + cutoffs of the C2 motif set on the device (cdf_kernel): timing of the batched call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bioinformatics_toolkit_b200 import _lib, synth
ctx = _lib.default_context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
mats = synth.synth_pwms(n, 12) if n <= 100 else synth.config_c4_motifs(n)
mo = _lib.DeviceMotifs(ctx, mats, (0.25,) * 4)
for rep in range(3):
    t0 = time.perf_counter(); th = mo.pvalue_to_score(1e-5); t1 = time.perf_counter()
    print("%d motifs: pvalue_to_score %.1f ms (kernel %.1f ms)" % (n, 1e3 * (t1 - t0), ctx.timing()["aux_kernel_ms"]))
