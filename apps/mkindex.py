#!/usr/bin/env python
"""mkindex OUT in1.fa ... — bioinformatics-toolkit-apps/app/MkIndex.hs.
mkindex --packed OUT in1.fa ... also writes OUT.packed: the 2-bit + invalid-base planes of every chromosome (0.375 bytes per
base), which whole-genome scans upload instead of the bytes (scan_motif --genome-wide uses it when it is there)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bioinformatics_toolkit_b200.seq_io import Genome, mkIndex, mkPackedIndex   # noqa: E402

if __name__ == "__main__":
    args = sys.argv[1:]
    packed = "--packed" in args
    args = [a for a in args if a != "--packed"]
    mkIndex(args[1:], args[0])
    if packed:
        with Genome(args[0]) as g:
            mkPackedIndex(g)
