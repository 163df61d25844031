"""Freezes oracle outputs as regression goldens (tests/golden/oracle_goldens.json).

These are NOT reference outputs (the Haskell reference cannot run here); they pin the oracle
against accidental change and give the GPU box known answers to compare with.  Floats are stored
as hex strings so comparisons are bit-exact.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)


def main():
    from oracle import oracle as O
    from bioinformatics_toolkit_b200 import synth
    mots = json.load(open(os.path.join(HERE, "ref_motifs_fasta.json")))["motifs"]
    meme = json.load(open(os.path.join(HERE, "ref_motifs_meme.json")))["motifs"]
    bg = O.DEF_BG
    out = {"note": "oracle outputs, frozen; floats as float.hex()", "fasta": [], "meme": [], "align": []}
    dna = synth.synth_dna(5000, 2)
    for m in mots:
        mat = np.array(m["rows"])
        xs, ps = O.score_cdf(bg, mat)
        th = O.pvalue_to_score(1e-4, bg, mat)
        pos, sc = O.find_tfbs(bg, mat, dna, 0.6 * O.optimal_score(bg, mat), True)
        out["fasta"].append({
            "name": m["name"], "pvalue_to_score_1e-4": th.hex(), "cdf_points": int(xs.size),
            "cdf_x_first": xs[0].hex(), "cdf_x_last": xs[-1].hex(), "cdf_p_first": ps[0].hex(), "cdf_p_last": ps[-1].hex(),
            "optimal_score": O.optimal_score(bg, mat).hex(),
            "sigma": [v.hex() for v in O.optimal_scores_suffix(bg, mat)],
            "tfbs_pos": pos.tolist(), "tfbs_sc": [v.hex() for v in sc],
            "max_match_sc": O.max_match_sc(bg, mat, dna).hex(),
        })
    for m in meme:
        mat = np.array(m["rows"])
        xs, ps = O.score_cdf(bg, mat)
        out["meme"].append({"name": m["name"], "pvalue_to_score_1e-5": O.pvalue_to_score(1e-5, bg, mat).hex(),
                            "cdf_points": int(xs.size)})
    mats = [np.array(m["rows"]) for m in mots] + [np.array(m["rows"]) for m in meme]
    for i in range(len(mats)):
        for j in range(len(mats)):
            d, (sd, p) = O.alignment(mats[i], mats[j])
            out["align"].append([i, j, d.hex(), bool(sd), int(p)])
    json.dump(out, open(os.path.join(HERE, "oracle_goldens.json"), "w"), indent=0)
    print("wrote oracle_goldens.json")


if __name__ == "__main__":
    main()
