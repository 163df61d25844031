"""Bisect of an abort at interpreter exit under MALLOC_PERTURB_: which entry point leaves something behind?  argv[1] = step."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bioinformatics_toolkit_b200 import _lib, synth
step = sys.argv[1]
ctx = _lib.default_context(0)
mats = synth.synth_pwms(20, 12)
BG = (0.25,) * 4
mo = _lib.DeviceMotifs(ctx, mats, BG)
dna = synth.synth_dna(2_000_000, 7, gc=0.41)
th = mo.pvalue_to_score(1e-4) if step != "nothing" else None
if step == "scan":
    h = _lib.Plan(mo, th).scan(_lib.DeviceSeq(ctx, dna)); print(len(h))
elif step == "scan_keep":
    seq = _lib.DeviceSeq(ctx, dna, keep_ascii=True); h = _lib.Plan(mo, th).scan(seq); print(len(h))
elif step == "iupac":
    seq = _lib.DeviceSeq(ctx, dna, keep_ascii=True); h = _lib.Plan(mo, th, skip_ambiguous=False).scan(seq); print(len(h))
elif step == "scores":
    seq = _lib.DeviceSeq(ctx, dna, keep_ascii=True); print(_lib.scores(seq, mo, 3).size)
elif step == "max":
    seq = _lib.DeviceSeq(ctx, dna, keep_ascii=True); print(_lib.max_match_sc(seq, mo)[:2])
elif step == "cdfs":
    print(len(mo.cutoff_cdfs(1e-4)[1]))
elif step == "affinity":
    h = _lib.Plan(mo, th).scan(_lib.DeviceSeq(ctx, dna)); _, cdfs = mo.cutoff_cdfs(1e-4)
    off = np.concatenate([[0], np.cumsum([c[0].size for c in cdfs])]).astype(np.int64)
    print(h.affinity(off, np.concatenate([c[0] for c in cdfs]), np.concatenate([c[1] for c in cdfs]))[:3])
elif step == "many":
    plan = _lib.Plan(mo, th); print(plan.scan_host_many([dna, dna[:500000]], consume=len))
print("done", step, flush=True)
