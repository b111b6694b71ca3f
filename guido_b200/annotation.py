"""GTF / GFF3 annotation tables with the conventions the reference inherits from pyranges.

guido reads annotations with ``pyranges.read_gtf`` / ``read_gff3`` (guido/locus.py:735-749):
columns ``Chromosome, Source, Feature, Start, End, Score, Strand, Frame`` plus one column per
attribute; ``Start`` is 0-based, ``End`` exclusive.  pyranges is absent from the target image, so
this module parses the files with pandas and applies the same renames
(GTF: gene_id->ID, exon_number->Exon; GFF3: Name->Exon).
"""
import os
import re

import pandas as pd

_BASE = ["Chromosome", "Source", "Feature", "Start", "End", "Score", "Strand", "Frame"]
_GTF_ATTR = re.compile(r'\s*([^\s;]+)\s+"?([^";]*)"?\s*;?')


def _parse_attrs(text, gff3):
    out = {}
    if gff3:
        for field in text.strip().strip(";").split(";"):
            if "=" in field:
                k, v = field.split("=", 1)
                out[k.strip()] = v.strip()
    else:
        for k, v in _GTF_ATTR.findall(text):
            if k not in out:
                out[k] = v
    return out


def read_annotation(path):
    """Parse a .gtf / .gff3 file into a DataFrame (pyranges-style coordinates, no renames)."""
    path = os.fspath(path)
    suffix = os.path.splitext(path)[1].lower()
    if suffix not in (".gtf", ".gff3", ".gff"):
        raise ValueError("Annotation file not recognised. Annotation file needs to be GFF3 or GTF formnat.")
    gff3 = suffix != ".gtf"
    rows = []
    with open(path) as fh:
        for line in fh:
            if not line.strip() or line.startswith("#"):
                continue
            cols = line.rstrip("\n").split("\t")
            if len(cols) < 9:
                continue
            row = {"Chromosome": cols[0], "Source": cols[1], "Feature": cols[2], "Start": int(cols[3]) - 1,
                   "End": int(cols[4]), "Score": cols[5], "Strand": cols[6], "Frame": cols[7]}
            row.update(_parse_attrs(cols[8], gff3))
            rows.append(row)
    df = pd.DataFrame(rows)
    if len(df) == 0:
        df = pd.DataFrame(columns=_BASE)
    return df


def renamed(df, path):
    suffix = os.path.splitext(os.fspath(path))[1].lower()
    if suffix in (".gff3", ".gff"):
        return df.rename(columns={"Name": "Exon"})
    return df.rename(columns={"gene_id": "ID", "exon_number": "Exon"})


def intersect(df, chromosome, start, end):
    """Features of ``chromosome`` overlapping [start, end) clipped to it (pyranges ``intersect``)."""
    if len(df) == 0:
        return df
    sel = df[(df["Chromosome"] == chromosome) & (df["Start"] < end) & (df["End"] > start)].copy()
    sel["Start"] = sel["Start"].clip(lower=start)
    sel["End"] = sel["End"].clip(upper=end)
    return sel.reset_index(drop=True)
