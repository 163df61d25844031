#!/usr/bin/env python
"""scan_motif GENOME MOTIF_MEME INPUT [-p 1e-5] — bioinformatics-toolkit-apps/app/MotifScan.hs.

withGenome -> readMEME -> map (mkCutoffMotif def p) -> streamBed (BED3) .| scanMotif .| stdout,
with the scan, the score CDFs and the cutoffs computed by the CUDA library.
INPUT = "-" (an addition): every chromosome of GENOME from end to end (bed_utils.scanGenome), read from GENOME.packed
(the 2-bit + mask planes, `mkindex --packed`) when that file exists, else from the bytes."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from bioinformatics_toolkit_b200.bed_utils import mkCutoffMotifs, scanGenome, scanMotif, streamBed3   # noqa: E402
from bioinformatics_toolkit_b200.motif import default_bkgd, readMEME                       # noqa: E402
from bioinformatics_toolkit_b200.seq_io import PackedGenome, withGenome                    # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser(prog="scan_motif")
    ap.add_argument("GENOME")
    ap.add_argument("MOTIF_MEME")
    ap.add_argument("INPUT")
    ap.add_argument("-p", "--p-value", type=float, default=1e-5, help="p-value cutoff. (default: 1e-5)")
    a = ap.parse_args(argv)

    def run(g):
        motifs = mkCutoffMotifs(default_bkgd, a.p_value, readMEME(a.MOTIF_MEME))
        if a.INPUT == "-":
            packed = PackedGenome(a.GENOME + ".packed") if os.path.exists(a.GENOME + ".packed") else None
            return scanGenome(g, motifs, out=sys.stdout.buffer, packed=packed)
        return scanMotif(g, motifs, streamBed3(a.INPUT), out=sys.stdout.buffer)
    withGenome(a.GENOME, run)


if __name__ == "__main__":
    main()
