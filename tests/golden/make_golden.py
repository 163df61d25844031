"""Regenerates tests/golden/*.json from the reference's own fixtures (run in the build container
only: /root/reference does not exist on the GPU box).

  python tests/golden/make_golden.py            # fixtures parsed from the reference test data
  python tests/golden/make_golden.py --oracle   # + freeze oracle outputs as regression goldens

ref_motifs_fasta.json / ref_motifs_meme.json: the PWMs of
bioinformatics-toolkit/tests/data/motifs.{fasta,meme}, parsed the way the reference parses them:
Bio.Utils.Misc.readDouble = bytestring-lexing's readExponential, restated in
bioinformatics_toolkit_b200.motif.read_double (integer mantissa, then ONE division by a power of ten in
Double — not correctly rounded: 15 of the 221 number tokens of these two files come out one ulp away
from strtod / Python float(); ref_motif_tokens.json keeps the raw tokens, tests/test_read_double.py the
list).  Parsing follows Data/Fasta.hs:37-51 + Motif.hs:329-330 (fasta) and Motif.hs:361-386 (MEME).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/bioinformatics-toolkit/tests/data"
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from bioinformatics_toolkit_b200.motif import read_double  # noqa: E402


def num(tok):
    return read_double(tok.encode("latin1"))


def parse_motif_fasta(text):
    recs, cur = [], None
    for line in text.split("\n"):
        if line == "":
            continue
        if line[0] == ">":
            cur = {"name": line[1:], "rows": []}
            recs.append(cur)
        elif cur is not None:
            cur["rows"].append([num(x) for x in line.split()])
    return recs


def parse_meme(text):
    out, st, cur = [], 0, None
    for x in text.split("\n"):
        if x.startswith("MOTIF"):
            st, cur = 1, {"name": x[6:], "rows": []}
            continue
        if st == 1:
            if x.startswith("letter-probability matrix:"):
                cur["nsites"] = int(x.split()[7])
                st = 2
        elif st == 2:
            y = x.lstrip(" ")
            if y == "" or y.startswith("URL"):
                out.append(cur)
                st, cur = 0, None
            else:
                cur["rows"].append([num(v) for v in y.split()])
    if st == 2:
        out.append(cur)
    return out


def main():
    with open(os.path.join(REF, "motifs.fasta")) as f:
        fasta = parse_motif_fasta(f.read())
    with open(os.path.join(REF, "motifs.meme")) as f:
        meme = parse_meme(f.read())
    json.dump({"source": "bioinformatics-toolkit/tests/data/motifs.fasta", "motifs": fasta},
              open(os.path.join(HERE, "ref_motifs_fasta.json"), "w"), indent=0)
    json.dump({"source": "bioinformatics-toolkit/tests/data/motifs.meme", "motifs": meme},
              open(os.path.join(HERE, "ref_motifs_meme.json"), "w"), indent=0)
    # the raw number tokens of both fixtures (for motif.read_double, the restatement of bytestring-lexing's readExponential)
    toks = set()
    for fn in ("motifs.fasta", "motifs.meme"):
        with open(os.path.join(REF, fn)) as f:
            for line in f.read().split("\n"):
                w = line.split()
                if len(w) == 4 and not line.startswith(("A ", ">", "MOTIF", "letter")):
                    toks.update(w)
    json.dump({"source": "number tokens of bioinformatics-toolkit/tests/data/motifs.{fasta,meme}", "tokens": sorted(toks)},
              open(os.path.join(HERE, "ref_motif_tokens.json"), "w"), indent=0)
    print("fasta motifs:", [(m["name"], len(m["rows"])) for m in fasta])
    print("meme motifs:", [(m["name"], len(m["rows"])) for m in meme])
    if "--oracle" in sys.argv:
        sys.path.insert(0, os.path.join(HERE, "..", ".."))
        from tests.golden import freeze_oracle
        freeze_oracle.main()


if __name__ == "__main__":
    main()
