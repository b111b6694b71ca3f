"""Convert the reference's CFD pickles (Doench et al. 2016 matrices, guido/data/*.pkl) into
guido_b200/data/cfd_tables.json.  Values are stored as float.hex() so they round-trip bit-exactly.
Run in the build container (needs /root/reference): python tools/make_cfd_tables.py
"""
import json
import os
import pickle

REF = os.environ.get("GUIDO_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "guido_b200", "data",
                   "cfd_tables.json")

mm = pickle.load(open(os.path.join(REF, "guido/data/mismatch_score.pkl"), "rb"))
pam = pickle.load(open(os.path.join(REF, "guido/data/PAM_scores.pkl"), "rb"))
with open(OUT, "w") as fh:
    json.dump({"source": "Doench et al. 2016 CFD matrices as shipped in nkran/guido guido/data/",
               "mismatch_scores": {k: float(v).hex() for k, v in sorted(mm.items())},
               "pam_scores": {k: float(v).hex() for k, v in sorted(pam.items())}}, fh, indent=0)
print(OUT, len(mm), len(pam))
