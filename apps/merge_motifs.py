#!/usr/bin/env python
"""merge_motifs {merge,dist,cluster} — bioinformatics-toolkit-apps/app/MergeMotifs.hs.

dist (158-169): name<TAB>name<TAB>distance for all i<j (or A x B); merge (171-201): iter | tree;
cluster (203-208).  All alignments run on the device in batched launches."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from bioinformatics_toolkit_b200 import _lib, merge as M                                    # noqa: E402
from bioinformatics_toolkit_b200.alignment import AlignFn                                    # noqa: E402
from bioinformatics_toolkit_b200.motif import (Motif, _shortest, default_bkgd, readFasta_,   # noqa: E402
                                               readMEME, toIUPAC, writeFasta, writeMEME)

GAP = {"linear": "linear", "quadratic": "quadratic", "cubic": "cubic", "exp": "exp"}
AVG = {"l1": "l1", "l2": "l2", "l3": "l3", "max": "max"}


def read_motif(fl):
    return readFasta_(fl) if fl.rsplit(".", 1)[-1] == "fasta" else readMEME(fl)


def write_motif(fl, motifs):
    if fl.rsplit(".", 1)[-1] == "fasta":
        writeFasta(fl, motifs)
    else:
        writeMEME(fl, motifs, default_bkgd)


def mk_align(a):
    if a.gap_mode not in GAP:
        raise SystemExit("Unknown gap mode")
    if a.avg_mode not in AVG:
        raise SystemExit("Unknown average mode")
    return AlignFn(GAP[a.gap_mode], a.gap, AVG[a.avg_mode])


def add_align(p):
    p.add_argument("-g", "--gap", type=float, default=0.05)
    p.add_argument("--gap_mode", default="exp")
    p.add_argument("--avg_mode", default="l1")


def main(argv=None):
    ap = argparse.ArgumentParser(prog="merge_motifs")
    sub = ap.add_subparsers(dest="cmd", required=True)
    p = sub.add_parser("merge")
    p.add_argument("INPUT"); p.add_argument("-m", "--mode", default="iter"); p.add_argument("-t", "--thres", type=float, default=0.2)
    add_align(p); p.add_argument("-o", "--output", default="merged_output.meme")
    p = sub.add_parser("dist")
    p.add_argument("-a", dest="INPUT_A", required=True); p.add_argument("-b", dest="inputB", default=None); add_align(p)
    p = sub.add_parser("cluster", add_help=False)       # the reference binds -h to --height
    p.add_argument("INPUT"); p.add_argument("-h", "--height", type=float, default=0.2); add_align(p)
    a = ap.parse_args(argv)
    align = mk_align(a)
    out = sys.stdout.buffer
    if a.cmd == "dist":
        ma = read_motif(a.INPUT_A)
        if a.inputB is None:
            d, _, _ = align.all_pairs([m._pwm for m in ma]) if len(ma) > 1 else (np.zeros(0), None, None)
            out.write(_lib.format_dist([m._name for m in ma], d))
        else:
            mb = read_motif(a.inputB)
            pw = [m._pwm for m in ma] + [m._pwm for m in mb]
            ia = np.repeat(np.arange(len(ma), dtype=np.int32), len(mb))
            ib = np.tile(np.arange(len(ma), len(ma) + len(mb), dtype=np.int32), len(ma))
            d, _, _ = align.pairs(pw, ia, ib)
            out.write(_lib.format_dist([m._name for m in ma] + [m._name for m in mb], d, ia, ib))
    elif a.cmd == "merge":
        motifs = read_motif(a.INPUT)
        print("Merging Mode: %s" % a.mode, file=sys.stderr)
        print("Read %d motifs" % len(motifs), file=sys.stderr)
        if a.mode == "tree":
            tree = M.buildTree(align, motifs)
            res = []
            for i, tr in enumerate(M.cutAt(tree, a.thres)):
                pwm = M.dilute(M.mergeTreeWeighted(align, tr))
                suffix = b"+".join(m._name for m in M.flatten(tr))
                res.append(Motif(("_%d_%s" % (i, toIUPAC(pwm))).encode() + b"(" + suffix + b")", pwm))
        else:
            # the reference computes `dilute` but writes the undiluted PWM (MergeMotifs.hs:194-196): kept
            res = [Motif(b"+".join(nm), pwm) for nm, pwm, ws in M.iterativeMerge(align, a.thres, motifs)]
        print("Write %d motifs" % len(res), file=sys.stderr)
        write_motif(a.output, res)
    else:
        motifs = read_motif(a.INPUT)
        for t in M.cutAt(M.buildTree(align, motifs), a.height):
            out.write(b"\t".join(m._name for m in M.flatten(t)) + b"\n")


if __name__ == "__main__":
    main()
