"""Config-4 chunk (64 Mbp): first-stage time of every 128-motif block on its own (the planner sorts motifs by length; a plan
made from one block's motifs reproduces that block), cold (L2 flushed) and warm (same scan repeated)."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bioinformatics_toolkit_b200 import _lib, synth
BG = (0.25,) * 4
ctx = _lib.default_context(0)
mats = synth.config_c4_motifs(800)
th = np.array([float.fromhex(v) for v in json.load(open("tests/golden/config_thresholds.json"))["c4"]])
dna = synth.synth_dna(int(sys.argv[1]) if len(sys.argv) > 1 else 64_000_000, 5, gc=0.41)
seq = _lib.DeviceSeq(ctx, dna)
order = sorted(range(800), key=lambda i: mats[i].shape[0])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(idx, label):
    mo = _lib.DeviceMotifs(ctx, [mats[i] for i in idx], BG)
    plan = _lib.Plan(mo, th[idx])
    info = plan.info()
    res = {}
    for mode in ("cold", "warm"):
        ts = []
        for rep in range(4):
            if mode == "cold":
                flush.fill_(1); torch.cuda.synchronize()
            h = plan.scan(seq); tm = ctx.timing(); ts.append(tm["prefilter_ms"]); nh = len(h); del h
        res[mode] = min(ts[1:])
    print("%-10s motifs %3d cols %2d..%2d sectors/pos %.4f surv/pos %.4f hits %8d  first stage cold %.3f ms warm %.3f ms" % (
        label, len(idx), min(mats[i].shape[0] for i in idx), max(mats[i].shape[0] for i in idx), info["table_sectors_per_position"],
        info["first_stage_survivors_per_position"], nh, res["cold"], res["warm"]), flush=True)
    plan.free(); mo.free()
for b in range(7):
    run(order[128 * b: 128 * b + 128], "block %d" % b)
run(list(range(800)), "all 7")
