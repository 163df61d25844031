"""p = 1e-5 cutoffs of the BASELINE configs' synthetic motifs, from the ORACLE's pValueToScore
(tests/golden/config_thresholds.json, floats as hex).  They are (a) known answers the device
scoreCDF / pValueToScore must reproduce bit for bit, and (b) the cutoffs the parity tests scan with.
Run here (build container):  python tests/golden/make_config_thresholds.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def main():
    from oracle import oracle as O
    from bioinformatics_toolkit_b200 import synth
    bg = O.DEF_BG
    out = {"note": "oracle pValueToScore 1e-5, uniform background; float.hex()"}
    out["c1"] = [O.pvalue_to_score(1e-5, bg, m).hex() for m in [synth.ctcf_like_pwm()]]
    out["c2"] = [O.pvalue_to_score(1e-5, bg, m).hex() for m in synth.synth_pwms(100, 12)]
    out["c4"] = [O.pvalue_to_score(1e-5, bg, m).hex() for m in synth.config_c4_motifs(800)]
    json.dump(out, open(os.path.join(HERE, "config_thresholds.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
