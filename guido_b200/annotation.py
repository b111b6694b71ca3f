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


def annotate_intervals(df, chromosomes, starts, ends, features=("CDS", "exon", "transcript", "gene"),
                       id_column=None):
    """Label query intervals with the annotation they fall into (off-target
    hits against the GTF -- the consumer north_star places after the search; the reference
    snapshot only builds the tabix index for it, guido/genome.py:145-148, and has no lookup).

    ``df``: a table from :func:`read_annotation` (0-based ``Start``, exclusive ``End``), renamed or
    not.  Queries are half-open ``[starts[i], ends[i])`` on ``chromosomes[i]``.  Returns a DataFrame
    with one row per query: ``feature`` = the most specific overlapping feature type in the order of
    ``features`` (``"intergenic"`` when none overlaps) and ``id`` = the ``ID`` / ``gene_id`` of the
    first (lowest ``Start``, then file order) overlapping record of that type ("" when none).

    Vectorised: per chromosome and feature type the records are sorted by ``Start`` and a running
    maximum of ``End`` says whether anything that starts before the query's end reaches its start;
    1e7 hits against 1e5 records take a few seconds of numpy, no Python loop over hits.
    """
    import numpy as np
    chromosomes = np.asarray(chromosomes, dtype=object)
    starts = np.asarray(starts, dtype=np.int64)
    ends = np.asarray(ends, dtype=np.int64)
    n = len(starts)
    feature = np.full(n, "intergenic", dtype=object)
    ident = np.full(n, "", dtype=object)
    if n == 0 or len(df) == 0:
        return pd.DataFrame({"feature": feature, "id": ident})
    if id_column is None:
        id_column = "ID" if "ID" in df.columns else "gene_id" if "gene_id" in df.columns else None
    done = np.zeros(n, bool)
    for chrom in pd.unique(chromosomes):
        q = np.nonzero(chromosomes == chrom)[0]
        sub = df[df["Chromosome"] == chrom]
        if len(sub) == 0:
            continue
        for ftype in features:
            recs = sub[sub["Feature"] == ftype]
            todo = q[~done[q]]
            if len(recs) == 0 or len(todo) == 0:
                continue
            order = np.argsort(recs["Start"].to_numpy(np.int64), kind="stable")
            rs = recs["Start"].to_numpy(np.int64)[order]
            re_ = recs["End"].to_numpy(np.int64)[order]
            ids = (recs[id_column].fillna("").astype(str).to_numpy(object)[order] if id_column else
                   np.full(len(recs), "", dtype=object))
            # records with Start < query end are rs[:k]; one of them overlaps iff its End > query start.
            # prefix maximum of End (and where it is attained first) answers "any" and names a witness;
            # the FIRST overlapping record is found by a second search on the prefix maxima.
            pmax = np.maximum.accumulate(re_)
            k = np.searchsorted(rs, ends[todo], side="left")          # number of records starting before the end
            has = k > 0
            hit = np.zeros(len(todo), bool)
            hit[has] = pmax[k[has] - 1] > starts[todo][has]
            if not hit.any():
                continue
            # first index j < k with End_j > start: pmax is non-decreasing, so the first j whose prefix
            # maximum exceeds start is the first record that does
            j = np.searchsorted(pmax, starts[todo][hit], side="right")
            sel = todo[hit]
            feature[sel] = ftype
            ident[sel] = ids[j]
            done[sel] = True
    return pd.DataFrame({"feature": feature, "id": ident})
