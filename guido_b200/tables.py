"""Columnar views for genome-scale work.

The reference materialises one Python object per guide and one dict per off-target; that is fine
for a locus (10^2-10^4 guides) and hopeless for a genome (3.4e7 NGG guides in 273 Mb, 1.1e7
off-target hits for 100 k guides).  These tables keep the kernels' output as numpy columns and
build the reference's records lazily, on demand, with exactly the values the per-locus API gives
(`tests/test_gpu_tables.py` checks that equivalence record for record).
"""
import numpy as np

from . import _native
from .guides import Guide, guide_geometry
from .helpers import load_cfd_scoring_matrix, rev_comp, rev_comp_fast
from .off_targets import decode_hits

_WEIGHTS = np.array([10, 5, 4, 3, 1], dtype=np.int64)  # calculate_ot_sum_score, off_targets.py:9-18


class GuideTable:
    """All gRNAs of a genome image for one PAM, as columns.

    Equivalent, contig by contig, to ``Locus(sequence=<contig>, name=<contig name>).find_guides(pam)``
    of the reference (i.e. ``locus_from_sequence``): same guides, same order (cut position, '+'
    before '-' on ties), same ids per contig.  Passing ``end=len(contig)`` explicitly makes the
    reference drop the contig's last base (``locus.py:298-307``); the table follows the default.
    Only 3-nt PAMs give guides (reference ``locus.py:316``).
    """

    def __init__(self, image, pam="NGG"):
        self.image, self.pam = image, pam
        P = len(pam)
        if P == 3:
            fw, rv = image.pam_scan_genome(pam, _native.SCAN_GUIDE)
        else:
            _native.pam_scan_seq("", pam)  # same KeyError for letters outside the IUPAC table
            fw = rv = np.zeros(0, np.uint32)
        cf, pf = image.locate(fw)
        cr, pr = image.locate(rv)
        contig = np.concatenate([cf, cr])
        strand_minus = np.concatenate([np.zeros(len(fw), bool), np.ones(len(rv), bool)])
        pam_pos = np.concatenate([pf, pr])                       # 0-based PAM start in the contig
        cut = np.where(strand_minus, pam_pos + P + 3, pam_pos - 3) + 1   # Locus start = 1
        order = np.lexsort((np.arange(len(cut)), cut, contig))  # stable: '+' first on ties
        self.contig = contig[order]
        self.strand_minus = strand_minus[order]
        self.pam_pos = pam_pos[order]
        self.absolute_cut_pos = cut[order]
        self.guide_start = np.where(self.strand_minus, self.pam_pos, self.pam_pos - 20) + 1
        self.guide_end = self.guide_start + 20 + P - 1
        # per-contig running index -> "gRNA-<k>" ids as find_guides assigns them
        first = np.searchsorted(self.contig, np.arange(len(image.contigs)), side="left")
        self.rank_in_contig = np.arange(len(cut)) - first[self.contig] + 1

    def __len__(self):
        return len(self.contig)

    def protospacers(self, index=None):
        """(n, 20) uint8 letters, guide orientation, straight from the image."""
        idx = np.arange(len(self)) if index is None else np.asarray(index)
        goff = np.array([c[2] for c in self.image.contigs], dtype=np.int64)
        start = goff[self.contig[idx]] + self.guide_start[idx] - 1
        w = self.image.fetch_windows(start, 20 + len(self.pam))
        comp = np.zeros(256, np.uint8)
        for a, b in zip(b"ACGTN", b"TGCAN"):
            comp[a] = b
        rc = comp[w[:, ::-1]]
        out = np.where(self.strand_minus[idx][:, None], rc, w)
        return np.ascontiguousarray(out[:, :20])

    def guide(self, i):
        """Materialise guide i as the reference's ``Guide`` record."""
        i = int(i)
        ci = int(self.contig[i])
        name, clen, _ = self.image.contigs[ci]
        p = int(self.pam_pos[i])
        # the geometry only looks at [p-99, p+P+99): fetch that slice and shift coordinates
        lo, hi = max(p - 110, 0), min(p + len(self.pam) + 110, clen)
        seq = self.image.fetch(ci, lo, hi - lo)
        strand = "-" if self.strand_minus[i] else "+"
        geo = guide_geometry(seq, p - lo, len(self.pam), strand, start=1 + lo,
                             rc_sequence=rev_comp_fast(seq) if strand == "-" else None)
        # the MMEJ window is cut from the whole contig (interval = contig), not from the slice
        cut0 = int(self.absolute_cut_pos[i]) - 1
        a, b = max(cut0 - 75, 0), min(cut0 + 75, clen)
        geo["locus_seq"] = self.image.fetch(ci, a, b - a)
        geo["relative_cut_pos"] = 75 if cut0 - 75 >= 0 else cut0
        g = Guide._from_geometry(geo, name, strand)
        g.id = f"gRNA-{int(self.rank_in_contig[i])}"
        return g

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self.guide(k) for k in range(*i.indices(len(self)))]
        return self.guide(i)

    # ---- exporters (helpers.py:66-92), straight from the columns -------------------------
    def to_dataframe(self, with_sequences=False):
        """One row per guide with the position columns of ``guides_to_dataframe`` (same names, same
        values); ``with_sequences`` adds ``guide_seq`` (protospacer + genomic PAM, fetched from the
        image).  Layers, MMEJ patterns and off-targets belong to materialised ``Guide`` records."""
        import pandas as pd
        names = np.array([c[0] for c in self.image.contigs], dtype=object)
        ids = np.char.add("gRNA-", self.rank_in_contig.astype(str))
        df = pd.DataFrame({
            "guide_chrom": names[self.contig], "guide_start": self.guide_start, "guide_end": self.guide_end,
            "guide_strand": np.where(self.strand_minus, "-", "+"), "absolute_cut_pos": self.absolute_cut_pos,
            "rank_score": 0.0, "rank": 0, "id": ids})
        if with_sequences and len(self):
            goff = np.array([c[2] for c in self.image.contigs], dtype=np.int64)
            w = self.image.fetch_windows(goff[self.contig] + self.guide_start - 1, 20 + len(self.pam))
            comp = np.zeros(256, np.uint8)
            for a, b in zip(b"ACGTN", b"TGCAN"):
                comp[a] = b
            w = np.where(self.strand_minus[:, None], comp[w[:, ::-1]], w)
            df["guide_seq"] = [r.tobytes().decode() for r in w]
        df.index = df["id"]
        return df

    def to_bed(self, file):
        """chrom, start, end, id, rank, strand -- the columns and layout of ``guides_to_bed``"""
        cols = ["guide_chrom", "guide_start", "guide_end", "id", "rank", "guide_strand"]
        self.to_dataframe().loc[:, cols].to_csv(file, sep="\t", header=False, index=False)


class OffTargetTable:
    """Off-target hits of a guide set as columns, with the reference's filters applied vectorised.

    ``guide_chrom`` / ``guide_start`` are per-guide arrays (names and 1-based starts) used for the
    reference's "another chromosome only" filter (``off_targets.py:199``).
    """

    def __init__(self, image, protos_u8, hits, raw_flags, pam, guide_chrom=None, guide_start=None):
        self.image, self.pam = image, pam
        P = len(pam)
        self.width = 20 + P
        letters = decode_hits(hits, self.width)
        contig_ix, offset = image.locate(hits["gpos"])
        names = np.array([c[0] for c in image.contigs], dtype=object)
        diff = letters[:, :20] != protos_u8[hits["guide"]]
        n_amb = sum(1 for c in pam if c not in "ACGT")
        keep = np.ones(len(hits), bool)
        if guide_chrom is not None:
            gc = np.asarray([str(c) for c in guide_chrom], dtype=object)[hits["guide"]]
            gs = np.asarray(guide_start, dtype=np.int64)[hits["guide"]]
            keep &= (gc != names[contig_ix]) & (gs != offset + 1)
        n_mm = diff.sum(axis=1) + n_amb
        keep &= n_mm > 0
        self.guide = hits["guide"][keep]
        self.chromosome = names[contig_ix][keep]
        self.start = offset[keep]
        self.strand_minus = (hits["hi"][keep] >> 31).astype(bool)
        self.letters = letters[keep]
        self.diff = diff[keep]
        self.n_mismatches = n_mm[keep]          # len(record["mismatches"]) incl. the PAM's N
        self.has_key = np.asarray(raw_flags, bool) if raw_flags is not None else None
        self.n_guides = len(protos_u8)
        self._protos = np.ascontiguousarray(protos_u8, dtype=np.uint8)

    def __len__(self):
        return len(self.guide)

    def ot_sum_score(self):
        """per-guide ``calculate_ot_sum_score`` (weights indexed by len(mismatches) - 1)"""
        w = _WEIGHTS[np.clip(self.n_mismatches - 1, 0, 4)]
        return np.bincount(self.guide, weights=w, minlength=self.n_guides).astype(np.int64)

    def counts_matrix(self):
        """(n_guides, 5) counts by len(mismatches) - 1, the rows of ``get_off_targets_string``"""
        m = np.zeros((self.n_guides, 5), np.int64)
        np.add.at(m, (self.guide, np.clip(self.n_mismatches - 1, 0, 4)), 1)
        return m

    def cfd_scores(self, guide_pams=None, mm_scores=None, pam_scores=None):
        """CFD score of every hit (``calculate_cfd_score``, off_targets.py:31-61), vectorised.

        The reference multiplies, position by position, the mismatch factors of the 20 protospacer
        positions and then the factor of the GUIDE's own PAM dinucleotide, in float64.  Here the
        same factors are taken from a (20, 4, 4) table (1.0 where the bases agree -- multiplying
        by 1.0 is exact) and multiplied column by column in the same order, so every score is
        bit-identical.  ``guide_pams``: the guides' own PAM strings (default: ``pam`` for every
        guide, which needs a PAM whose last two letters are A/C/G/T).
        """
        if mm_scores is None or pam_scores is None:
            mm_scores, pam_scores = load_cfd_scoring_matrix()
        factor = np.ones((20, 4, 4), dtype=np.float64)     # [position, guide base, off-target base]
        for i in range(20):
            for a, wt in enumerate("ACGU"):
                for b, base in enumerate("ACGU"):
                    if wt != base:
                        factor[i, a, b] = mm_scores["r" + wt + ":d" + rev_comp(base, rna=True) + "," + str(i + 1)]
        code = np.zeros(256, np.intp)
        for k, c in enumerate(b"ACGT"):
            code[c] = k
        wt = code[self._protos[self.guide]]                  # (n, 20)
        ot = code[self.letters[:, :20]]
        score = np.ones(len(self), dtype=np.float64)
        for i in range(20):
            score = score * factor[i, wt[:, i], ot[:, i]]
        if guide_pams is None:
            pam_factor = np.full(self.n_guides, pam_scores[self.pam[1:]], dtype=np.float64)
        else:
            pam_factor = np.array([pam_scores[p[1:]] for p in guide_pams], dtype=np.float64)
        return score * pam_factor[self.guide]

    def cfd_layers(self, guide_pams=None):
        """per-guide (mean, max, sum) of the CFD scores -- the ``ot_cfd_score_*`` layers of
        ``Locus.find_off_targets`` (NaN for guides without a hit).  Sums run left to right over a
        guide's hits; numpy's pairwise ``sum()`` in the reference can differ in the last ulp."""
        cfd = self.cfd_scores(guide_pams)
        order = np.argsort(self.guide, kind="stable")
        g, c = self.guide[order], cfd[order]
        n = np.bincount(g, minlength=self.n_guides)
        total = np.zeros(self.n_guides, np.float64)
        best = np.full(self.n_guides, np.nan)
        if len(c):
            starts = np.concatenate(([0], np.cumsum(n)[:-1]))
            has = n > 0
            total[has] = np.add.reduceat(c, starts[has])
            best[has] = np.maximum.reduceat(c, starts[has])
        with np.errstate(invalid="ignore", divide="ignore"):
            mean = np.where(n > 0, total / np.maximum(n, 1), np.nan)
        return mean, best, np.where(n > 0, total, np.nan)

    def annotate(self, annotation, features=("CDS", "exon", "transcript", "gene")):
        """Per-hit annotation (DataFrame: guide, chromosome, start, strand, feature, id): the most
        specific feature of ``annotation`` -- a table from ``guido_b200.annotation.read_annotation``
        or ``Genome.annotation`` -- that the 23-nt (20 + PAM) target overlaps, ``"intergenic"``
        otherwise.  SURVEY §8(f) row 4; vectorised, see ``annotation.annotate_intervals``."""
        import pandas as pd
        from .annotation import annotate_intervals
        lab = annotate_intervals(annotation, self.chromosome, self.start, self.start + self.width, features)
        return pd.DataFrame({"guide": self.guide, "chromosome": self.chromosome, "start": self.start,
                             "strand": np.where(self.strand_minus, "-", "+"),
                             "feature": lab["feature"].to_numpy(), "id": lab["id"].to_numpy()})

    def records(self, ix):
        """the reference's list of off-target dicts for guide ``ix`` (no ``cfd_score``)"""
        P = len(self.pam)
        amb = [k for k, c in enumerate(self.pam) if c not in "ACGT"]
        out = []
        for h in np.nonzero(self.guide == ix)[0].tolist():
            window = self.letters[h].tobytes().decode()
            mm = {}
            for k in reversed(amb):
                mm[21 + k] = window[20 + k]
            for i in np.nonzero(self.diff[h])[0][::-1].tolist():
                mm[i + 1] = window[i]
            out.append({"ix": int(ix), "mismatches": mm,
                        "mismatches_string": "".join(window[i] if (i + 1) in mm else "." for i in range(20 + P)),
                        "chromosome": self.chromosome[h], "start": int(self.start[h]),
                        "strand": "-" if self.strand_minus[h] else "+", "seq": window})
        return out


def offtarget_table(image, protos, pam="NGG", core_length=10, core_mismatches=2, total_mismatches=4,
                    guide_chrom=None, guide_start=None, algo=_native.OT_AUTO):
    """Search and wrap: ``protos`` is a list of 20-nt strings or an (n, 20) uint8 array."""
    if not isinstance(protos, np.ndarray):
        protos = np.frombuffer("".join(protos).encode(), np.uint8).reshape(-1, 20)
    hits, flags = image.offtarget_search(protos, pam=pam, core_length=core_length,
                                         core_mismatches=core_mismatches,
                                         total_mismatches=total_mismatches, algo=algo)
    return OffTargetTable(image, protos, hits, flags, pam, guide_chrom, guide_start)
