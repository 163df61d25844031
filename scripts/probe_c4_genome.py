"""Config 4 on one GPU, whole genome generated once: ms per genome for 1..4 scans in flight (PWMSCAN_LANES), resident
and from host bytes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bioinformatics_toolkit_b200 import _lib
ctx = _lib.default_context(0)
t0 = time.perf_counter()
seqs_host, mats = bench.workload("c4", 0, 1)
print("genome synthesised in %.1f s (%d items)" % (time.perf_counter() - t0, len(seqs_host)), flush=True)
mo = _lib.DeviceMotifs(ctx, mats, bench.BG)
th = mo.pvalue_to_score(1e-5)
plan = _lib.Plan(mo, th)
pinned = [torch.from_numpy(a).pin_memory() for _, a, _ in seqs_host]
arrays = [p.numpy() for p in pinned]
dev = [_lib.DeviceSeq(ctx, a) for a in arrays]
units = sum(own for _, _, own in seqs_host) * len(mats)
for lanes in (sys.argv[1:] or ["1", "2", "3", "4"]):
    os.environ["PWMSCAN_LANES"] = lanes
    for name, fn in (("resident", lambda: plan.scan_many(dev, consume=len)), ("host", lambda: plan.scan_host_many(arrays, consume=len))):
        fn()
        ts = []
        for _ in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter(); n = sum(fn()); ts.append(time.perf_counter() - t0)
        tm = ctx.timing()
        print("lanes %s %-8s %.1f ms/genome (min %.1f) = %.0f Gbp*motifs/s | hits %d | %s" % (lanes, name, 1e3 * np.mean(ts), 1e3 * min(ts), units / np.mean(ts) / 1e9, n,
              {k: round(float(v), 1) for k, v in tm.items() if k.endswith("_ms") or k in ("d2h_bytes", "kernel_launches")}), flush=True)
