#!/usr/bin/env python
"""mkindex OUT in1.fa ... — bioinformatics-toolkit-apps/app/MkIndex.hs."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bioinformatics_toolkit_b200.seq_io import mkIndex   # noqa: E402

if __name__ == "__main__":
    mkIndex(sys.argv[2:], sys.argv[1])
