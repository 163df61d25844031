#!/usr/bin/env python
"""BASELINE config 3: the scan_motif app — 50 000 synthetic 500-bp BED peaks x 800 JASPAR-sized PWMs
over an indexed genome file (hg38-shaped chromosomes scaled by --scale), p < 1e-5; BED6 to a file.
Prints one JSON line with the phase timings and checks a sample of regions against the oracle app.
"""
import argparse
import io
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

from bioinformatics_toolkit_b200 import _lib, synth                                            # noqa: E402
from bioinformatics_toolkit_b200 import motif as Mo                                            # noqa: E402
from bioinformatics_toolkit_b200.bed_utils import mkCutoffMotifs, scanMotif, streamBed3       # noqa: E402
from bioinformatics_toolkit_b200.seq_io import openGenome, write_genome                       # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.1)
    ap.add_argument("--regions", type=int, default=50000)
    ap.add_argument("--check-regions", type=int, default=150)
    a = ap.parse_args()
    tmp = tempfile.mkdtemp()
    t0 = time.perf_counter()
    chroms = []
    for i, (nm, L) in enumerate(synth.HG38_CHROM_SIZES):
        L = max(2000, int(L * a.scale))
        runs = [(0, min(1000, L)), (L - 1000, 1000), (int(L * 0.45), int(L * 0.03))]
        chroms.append((nm.encode(), synth.synth_dna(L, 4 + 100 * i, gc=0.41, n_runs=runs, lower_frac=0.3)))
    gpath = os.path.join(tmp, "genome.idx")
    write_genome(chroms, gpath)
    rng = np.random.Generator(np.random.PCG64(3))
    sizes = np.array([len(c[1]) for c in chroms], dtype=np.float64)
    which = rng.choice(len(chroms), size=a.regions, p=sizes / sizes.sum())
    beds = []
    for c in which:
        s = int(rng.integers(0, len(chroms[c][1]) - 500))
        beds.append((chroms[c][0], s, s + 500))
    bpath = os.path.join(tmp, "peaks.bed")
    with open(bpath, "wb") as f:
        for c, s, e in beds:
            f.write(b"%s\t%d\t%d\n" % (c, s, e))
    mats = synth.config_c4_motifs(800)
    names = [b"MA%04d.1 synthetic_%d" % (i, i) for i in range(len(mats))]
    mpath = os.path.join(tmp, "motifs.meme")
    Mo.writeMEME(mpath, [Mo.Motif(n, Mo.PWM(1, m)) for n, m in zip(names, mats)])
    t_setup = time.perf_counter() - t0

    ctx = _lib.default_context(0)
    t0 = time.perf_counter()
    motifs = Mo.readMEME(mpath)
    t_read = time.perf_counter() - t0
    t0 = time.perf_counter()
    cms = mkCutoffMotifs(Mo.default_bkgd, 1e-5, motifs, ctx)
    t_cut = time.perf_counter() - t0
    g = openGenome(gpath)
    out = io.BytesIO()
    scanMotif(g, cms, [beds[0]], out=io.BytesIO(), warn=None, ctx=ctx)      # warm-up (plan, pools)
    t0 = time.perf_counter()
    n = scanMotif(g, cms, streamBed3(bpath), out=out, warn=None, ctx=ctx)
    t_scan = time.perf_counter() - t0
    t0 = time.perf_counter()
    scanMotif(g, cms, streamBed3(bpath), out=io.BytesIO(), warn=None, ctx=ctx)
    t_scan2 = time.perf_counter() - t0
    lines = out.getvalue().split(b"\n")[:-1]
    # parity on a sample of regions (the oracle-side restatement of mkCutoffMotif + scanMotif + toLine)
    from oracle import app_oracle
    parsed = [Mo.PWM(None, m._pwm._mat) for m in motifs]
    k = a.check_regions
    t0 = time.perf_counter()
    exp = app_oracle.scan_motif_lines(dict(chroms), [m._name for m in motifs], [p._mat for p in parsed], beds[:k], p=1e-5)
    t_oracle = time.perf_counter() - t0
    sub = io.BytesIO()
    scanMotif(g, cms, beds[:k], out=sub, warn=None, ctx=ctx)
    got = sub.getvalue().split(b"\n")[:-1]
    units = a.regions * 500 * len(mats)
    print(json.dumps({"config": "C3 scan_motif app: %d x 500-bp BED regions x %d PWMs, genome scale %.2f" % (a.regions, len(mats), a.scale),
                      "hits": n, "lines": len(lines), "setup_s": t_setup, "read_meme_s": t_read, "mk_cutoff_motifs_s": t_cut,
                      "scan_motif_s": t_scan, "scan_motif_repeat_s": t_scan2, "gbp_motifs_per_s_scan": units / t_scan / 1e9,
                      "parity_sample_regions": k, "parity_sample_lines": len(exp), "parity_ok": got == exp,
                      "oracle_sample_s": t_oracle, "oracle_gbp_motifs_per_s_1core": k * 500 * len(mats) / t_oracle / 1e9}))
    assert got == exp, "scan_motif output differs from the oracle app on the sample"


if __name__ == "__main__":
    main()
