"""Small host-side helpers that the hot path shares (mirror of guido/helpers.py's public names).

``rev_comp`` keeps the reference's exact behaviour (guido/helpers.py:19-26): every letter outside
the complement table becomes ``N``; with ``rna=True`` the table also knows ``U`` and ``-`` but not
``N``.  The CFD tables are the Doench et al. 2016 matrices the reference ships as pickles
(guido/data/*.pkl); here they live as JSON (``data/cfd_tables.json``, written by
``tools/make_cfd_tables.py`` with exact float round-trip).
"""
import json
import os
from shutil import which

_DNA = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
_RNA = {"A": "T", "C": "G", "G": "C", "T": "A", "U": "A", "-": "-"}

# bytes.translate tables: one C call reverse-complements a whole interval
_DNA_TABLE = bytes(ord(_DNA.get(chr(i), "N")) for i in range(256))

EXCLUDED_ATTRS = ["locus_seq", "guide_long_seq", "relative_cut_pos", "mmej_patterns", "off_targets",
                  "_layers"]


def rev_comp(seq, rna=False):
    table = _RNA if rna else _DNA
    return "".join(table.get(base, "N") for base in reversed(seq))


def rev_comp_fast(seq):
    """Same result as ``rev_comp(seq)`` for ASCII input, via ``bytes.translate``."""
    return seq.encode("latin-1", "replace").translate(_DNA_TABLE)[::-1].decode("latin-1")


def is_tool(name):
    """Whether ``name`` is an executable on PATH."""
    return which(name) is not None


_cfd_cache = None


def load_cfd_scoring_matrix():
    """(mm_scores, pam_scores) dicts with the reference's keys ('rU:dT,12', 'GG', ...)."""
    global _cfd_cache
    if _cfd_cache is None:
        path = os.path.join(os.path.dirname(os.path.realpath(__file__)), "data", "cfd_tables.json")
        with open(path) as fh:
            raw = json.load(fh)
        _cfd_cache = ({k: float.fromhex(v) for k, v in raw["mismatch_scores"].items()},
                      {k: float.fromhex(v) for k, v in raw["pam_scores"].items()})
    return dict(_cfd_cache[0]), dict(_cfd_cache[1])


def _guide_instance_to_dict(guide):
    d = {k: v for k, v in guide.__dict__.items() if k not in EXCLUDED_ATTRS}
    for name, value in guide.layers.items():
        d[name] = value.mean() if hasattr(value, "mean") and hasattr(value, "shape") else value
    return d


def _guides_to_dataframe(guides):
    import pandas as pd
    df = pd.DataFrame([_guide_instance_to_dict(g) for g in guides])
    if len(df):
        df.index = df["id"]
    return df


def _guides_to_csv(guides, file):
    _guides_to_dataframe(guides).to_csv(file)


def _guides_to_bed(guides, file):
    cols = ["guide_chrom", "guide_start", "guide_end", "id", "rank", "guide_strand"]
    _guides_to_dataframe(guides).loc[:, cols].to_csv(file, sep="\t", header=False, index=False)
