#!/usr/bin/env python
"""scan_motif GENOME MOTIF_MEME INPUT [-p 1e-5] — bioinformatics-toolkit-apps/app/MotifScan.hs.

withGenome -> readMEME -> map (mkCutoffMotif def p) -> streamBed (BED3) .| scanMotif .| stdout,
with the scan, the score CDFs and the cutoffs computed by the CUDA library."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from bioinformatics_toolkit_b200.bed_utils import mkCutoffMotifs, scanMotif, streamBed3   # noqa: E402
from bioinformatics_toolkit_b200.motif import default_bkgd, readMEME                       # noqa: E402
from bioinformatics_toolkit_b200.seq_io import withGenome                                  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser(prog="scan_motif")
    ap.add_argument("GENOME")
    ap.add_argument("MOTIF_MEME")
    ap.add_argument("INPUT")
    ap.add_argument("-p", "--p-value", type=float, default=1e-5, help="p-value cutoff. (default: 1e-5)")
    a = ap.parse_args(argv)

    def run(g):
        motifs = mkCutoffMotifs(default_bkgd, a.p_value, readMEME(a.MOTIF_MEME))
        return scanMotif(g, motifs, streamBed3(a.INPUT), out=sys.stdout.buffer)
    withGenome(a.GENOME, run)


if __name__ == "__main__":
    main()
