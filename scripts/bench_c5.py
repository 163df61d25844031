#!/usr/bin/env python
"""BASELINE config 5: mergeMotifs all-vs-all PWM alignment for 2 000 synthetic motifs
(alignmentBy jsd (expPenal 0.05) l1), plus the batched scoreCDF / pValueToScore of 800 motifs.
Prints one JSON line; a sample of pairs is checked against the oracle."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from bioinformatics_toolkit_b200 import _lib, synth   # noqa: E402

BG = (0.25, 0.25, 0.25, 0.25)


def main():
    from oracle import oracle as O
    ctx = _lib.default_context(0)
    mats = synth.config_c5_motifs(2000)
    mo = _lib.DeviceMotifs(ctx, mats, BG)
    mo.align_all_pairs()                                   # warm-up
    t0 = time.perf_counter()
    d, sd, pos = mo.align_all_pairs()
    t_all = time.perf_counter() - t0
    tm = ctx.timing()
    M = len(mats)
    npair = M * (M - 1) // 2
    # oracle sample (one core) for the CPU column and parity
    rng = np.random.default_rng(1)
    rows = rng.choice(M - 1, size=6, replace=False)
    t0 = time.perf_counter()
    nchk, ok = 0, True
    for i in rows:
        base = i * M - i * (i + 1) // 2
        for j in range(i + 1, min(M, i + 1 + 500)):
            rd, (rs, rp) = O.alignment(mats[i], mats[j])
            k = base + (j - i - 1)
            ok &= (bool(sd[k]), int(pos[k])) == (rs, rp) and abs(d[k] - rd) <= 1e-12 * max(abs(rd), 1e-300)
            nchk += 1
    t_or = time.perf_counter() - t0
    # CDF batch
    m800 = synth.config_c4_motifs(800)
    mo8 = _lib.DeviceMotifs(ctx, m800, BG)
    mo8.pvalue_to_score(1e-5)
    t0 = time.perf_counter()
    th = mo8.pvalue_to_score(1e-5)
    t_cdf = time.perf_counter() - t0
    cdf_kernel_ms = ctx.timing()["aux_kernel_ms"]
    t0 = time.perf_counter()
    for m in m800[:20]:
        O.pvalue_to_score(1e-5, BG, m)
    t_cdf_or = (time.perf_counter() - t0) / 20
    print(json.dumps({"config": "C5 all-vs-all alignment of %d PWMs (%d pairs), jsd/expPenal 0.05/l1" % (M, npair),
                      "align_call_s": t_all, "align_kernel_ms": tm["aux_kernel_ms"], "pairs_per_s": npair / t_all,
                      "host_reevaluated_pairs": int(tm["aux_host_fallbacks"]), "parity_pairs_checked": nchk, "parity_ok": bool(ok),
                      "oracle_pairs_per_s_1core": nchk / t_or,
                      "cdf_800_motifs_call_s": t_cdf, "cdf_kernel_ms": cdf_kernel_ms, "oracle_cdf_s_per_motif_1core": t_cdf_or}))
    assert ok


if __name__ == "__main__":
    main()
