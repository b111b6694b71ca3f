"""``Locus`` and the locus factories (drop-in for guido.locus, reference guido/locus.py).

Public names, signatures, coordinate conventions and returned records follow the reference; the
PAM scan (``re.finditer`` x2 at locus.py:182-183), the MMEJ enumeration and the off-target search
are delegated to the GPU through the C-ABI.  ``find_guides`` scans the locus ONCE on the device and
derives every interval's matches from that single hit list (a look-ahead regex match inside a
substring is the same match inside the whole string, as long as it fits the substring).

Out of scope here: data layers / variation statistics / TOPSIS ranking
(locus.py:409-688) and the Jinja report -- they work on <=10^3 guides and need scikit-allel / mcdm.
"""
import numpy as np
import pandas as pd

from . import _native
from . import mmej as _mmej
from .annotation import intersect, read_annotation, renamed
from .guides import Guide, guide_geometry
from .helpers import (_guides_to_bed, _guides_to_csv, _guides_to_dataframe, load_cfd_scoring_matrix,
                      rev_comp_fast)
from .off_targets import get_off_targets_string, run_bowtie
from .sequence import FastaFile, Sequence, is_sequence_like


class Locus:
    def __init__(self, sequence, name=None, start=1, end=None, genome=None, annotation=None, **kwargs):
        """A genomic locus in which gRNAs are searched.

        sequence: ``Sequence`` record (name/seq/start/end) or plain ``str``; name: contig name;
        start/end: 1-based coordinates of the locus; genome: ``Genome`` used by default for the
        off-target search; annotation: DataFrame of features (pyranges-style columns).
        """
        self.sequence = sequence
        self.chromosome = name
        self.start = start
        self.intervals = []
        self.genome = genome
        self.annotation = annotation
        self.pam = ""
        self.guides = []
        self._layers = {}
        # (contig name, 0-based start) when the sequence was cut from `genome` by one of the factory
        # functions: find_guides then scans the genome image resident in HBM instead of uploading
        # the locus sequence again
        self._image_region = None

        if is_sequence_like(sequence):
            self.start = sequence.start
            self.end = sequence.end
            self.chromosome = sequence.name
        else:
            self.sequence = Sequence(seq=sequence, start=start, end=end, name=name, **kwargs)
            if end and isinstance(end, int):
                self.end = end
            elif not end:
                self.end = self.start + len(self.sequence.seq)  # one past the last base, as upstream
        self.length = len(self.sequence)

    def __repr__(self):
        return f"Locus({self.chromosome}:{self.start}-{self.end})"

    def to_dict(self):
        return self.__dict__

    def guide(self, ix):
        """Guide by position (int) or by id ("gRNA-3")."""
        if isinstance(ix, str):
            for g in self.guides:
                if g.id == ix:
                    return g
            raise ValueError("Provided gRNA index is not valid.")
        if isinstance(ix, int):
            return self.guides[ix]
        raise ValueError("Provided gRNA index is not valid.")

    @property
    def layers(self):
        return self._layers

    # ------------------------------------------------------------------ intervals
    def _flatten_intervals(self, intervals):
        """Union of adjacent / overlapping [start, end] pairs, clamped to the locus (locus.py:144-156;
        only the first interval of a merged group is clamped, as upstream)."""
        merged = []
        for start, end in intervals:
            if merged and merged[-1][1] >= start - 1:
                merged[-1][1] = max(merged[-1][1], end)
                continue
            merged.append([max(start, self.start), min(end, self.end)])
        return merged

    # ------------------------------------------------------------------ PAM scan -> guides
    def _guides_from_pam_positions(self, fwd_seq, start, pam, pams_fw, pams_rv, keep=None):
        """Build Guide records for PAM starts (relative to ``fwd_seq``), forward strand first.
        ``keep`` optionally pre-filters (strand, position) pairs that find_guides would drop anyway."""
        P = len(pam)
        n = len(fwd_seq)
        rc_seq = rev_comp_fast(fwd_seq) if len(pams_rv) else None
        chrom = self.chromosome
        out = []
        make = Guide._from_fields
        for strand, positions, rc in (("+", pams_fw, None), ("-", pams_rv, rc_seq)):
            minus = strand == "-"
            for p in positions:
                if keep is not None and not keep(strand, p):
                    continue
                if p < 24:
                    # slices that can wrap around (guides.py:69-76 quirk): the general routine
                    geo = guide_geometry(fwd_seq, p, P, strand, start=start, rc_sequence=rc)
                    if "N" not in geo["guide_seq"]:
                        out.append(Guide._from_geometry(geo, chrom, strand))
                    continue
                # bulk path: guide_geometry inlined for p >= 24 (no negative slice index anywhere)
                if minus:
                    cut, rs, re = p + P + 3, p, p + P + 20
                    guide_seq = rc[max(n - re, 0):n - rs]
                    if "N" in guide_seq:
                        continue
                    long_seq = rc[max(n - re - 4, 0):n - rs + 3]
                else:
                    cut, rs, re = p - 3, p - 20, p + P
                    guide_seq = fwd_seq[rs:re]
                    if "N" in guide_seq:
                        continue
                    long_seq = fwd_seq[rs - 4:re + 3]
                lo = cut - 75
                if lo < 0:
                    window, rel_cut = fwd_seq[0:cut + 75], cut
                else:
                    window, rel_cut = fwd_seq[lo:cut + 75], 75
                out.append(make(window, guide_seq, long_seq, P, chrom, start + rs, start + re - 1,
                                strand, rel_cut, start + cut))
        return out

    @staticmethod
    def _long_pam(pam):
        """A PAM of more than 8 letters does not fit the device matcher -- and cannot give a guide anyway:
        the reference keeps a match only if its guide is 23 letters long, i.e. for 3-letter PAMs
        (locus.py:316).  Its letters are still checked like the reference's iupac_dict lookup
        (KeyError for a letter outside the table, locus.py:178); the result is "no match"."""
        if len(pam) <= 8:
            return False
        for c in pam:
            _native.pam_scan_seq("", c)   # raises KeyError for a letter outside the IUPAC table
        return True

    def _find_guides_in_interval(self, sequence, start, pam):
        """Every guide whose PAM lies in ``sequence`` (coordinate of sequence[0] = ``start``):
        forward-strand guides by ascending PAM position, then reverse-strand ones; guides whose
        23-mer holds an N are dropped (locus.py:158-211).  The PAM scan runs on the GPU."""
        fwd_seq = sequence.upper()
        if self._long_pam(pam):
            return []
        pams_fw, pams_rv = _native.pam_scan_seq(fwd_seq, pam)
        return self._guides_from_pam_positions(fwd_seq, start, pam, pams_fw.tolist(), pams_rv.tolist())

    def _scan_locus(self, upper, pam):
        """PAM starts of both strands relative to the locus start.  A locus cut from a built genome
        is scanned in place on the genome image in HBM (``guido_pam_scan_region``: no sequence
        traffic at all); any other sequence goes to the device for the call (``guido_pam_scan_seq``)."""
        if self._long_pam(pam):
            return np.zeros(0, np.int32), np.zeros(0, np.int32)
        region = self._image_region
        if region is not None and self.genome is not None and getattr(self.genome, "bowtie_index", None):
            try:
                img = self.genome.device_image()
                ci = img.contig_index(region[0])
                n = len(upper)
                if 0 <= region[1] and region[1] + n <= img.contigs[ci][1]:
                    k = min(n, 96)  # the resident image must be the sequence the locus holds
                    if img.fetch(ci, region[1], k) == upper[:k] and img.fetch(ci, region[1] + n - k, k) == upper[n - k:]:
                        return img.pam_scan_region(ci, region[1], region[1] + n, pam, _native.SCAN_REGEX)
            except (KeyError, OSError, ValueError, RuntimeError):
                # no usable image for this genome (built with bowtie_ignore, file moved, ...): the
                # sequence itself goes to the device -- still the GPU, which fails loudly if there is none
                pass
        return _native.pam_scan_seq(upper, pam)

    def find_guides(self, pam="NGG", min_flanking_length=0, selected_features="all"):
        """Find gRNAs in the locus (both strands), sorted by cut position and named gRNA-1..n.

        pam: IUPAC PAM (3' of a 20-nt protospacer); min_flanking_length: flank around each interval
        that is searched but in which cut sites are ignored; selected_features: restrict the search
        to annotation features of that type (``"all"`` = whole locus).
        """
        self.pam = pam
        self.guides = []
        self.intervals = [[self.start, self.end]]

        if "all" not in selected_features and isinstance(self.annotation, pd.DataFrame):
            ann = self.annotation
            feature_ok = (ann["Feature"].isin(list(selected_features)) if isinstance(selected_features, (list, set, tuple))
                          else ann["Feature"] == selected_features)
            inside = (((ann["Start"] >= self.start) & (ann["Start"] <= self.end))
                      | ((ann["End"] >= self.start) & (ann["End"] <= self.end)))
            rows = ann[feature_ok & (ann["Chromosome"] == self.chromosome) & inside].sort_values("Start")
            if len(rows) > 0:
                self.intervals = self._flatten_intervals(
                    [[s, e] for s, e in rows[["Start", "End"]].values])

        whole = self.sequence.seq
        n = len(whole)
        P = len(pam)
        upper = whole.upper()
        # one device scan for the whole locus; every interval slices the sorted hit lists
        all_fw, all_rv = self._scan_locus(upper, pam)

        for interval_start, interval_end in self.intervals:
            a = max(interval_start - min_flanking_length - self.start, 0)
            b = min(interval_end + min_flanking_length - self.start, n)
            seq_start = interval_start - min_flanking_length  # NOT clamped, as upstream (locus.py:309)
            sub = upper[a:b]
            m = len(sub)

            def rel(arr, minus):
                if m < P:
                    return []
                lo = np.searchsorted(arr, a, side="left")
                hi = np.searchsorted(arr, a + m - P, side="right")
                p = arr[lo:hi].astype(np.int64) - a
                # exactly the post-filter of locus.py:313-319 (23-nt guide, cut inside the
                # interval), evaluated on the position column before any record is built
                if minus:
                    cut, whole_guide = seq_start + p + P + 3, p + P + 20 <= m
                else:
                    cut, whole_guide = seq_start + p - 3, p >= 20
                return p[whole_guide & (cut >= interval_start) & (cut <= interval_end)].tolist()

            found = self._guides_from_pam_positions(sub, seq_start, pam, rel(all_fw, False), rel(all_rv, True))
            self.guides.extend(g for g in found
                               if len(g.guide_seq) == 23 and interval_start <= g.absolute_cut_pos <= interval_end)

        ordered = sorted(self.guides, key=lambda g: g.absolute_cut_pos)
        for i, g in enumerate(ordered):
            g.id = f"gRNA-{i + 1}"
        self.guides = ordered
        return ordered

    # ------------------------------------------------------------------ MMEJ
    def simulate_end_joining(self, n_patterns=5, length_weight=20):
        """MMEJ deletion patterns for every gRNA -- one batched GPU call for the whole locus."""
        if len(self.guides) == 0:
            raise ValueError("No gRNAs saved yet.")
        todo = [g for g in self.guides if "N" not in g.locus_seq]
        # only the n_patterns winners are materialised as dicts (their scores decide, and those
        # are evaluated for every kept candidate, bit-identically, in numpy)
        results = _mmej.top_patterns_batch([g.relative_cut_pos for g in todo],
                                           [g.locus_seq for g in todo], length_weight, n_patterns)
        by_guide = {id(g): r for g, r in zip(todo, results)}
        for g in self.guides:  # in order, so an upstream-style TypeError leaves later guides untouched
            if id(g) in by_guide:
                g._finish_end_joining(by_guide[id(g)], n_patterns)
            else:
                g._skip_end_joining()

    # ------------------------------------------------------------------ off-targets
    def find_off_targets(self, external_genome=None, **kwargs):
        """Off-targets of every gRNA in ``external_genome`` (or the locus' own genome); attaches the
        hits, ``ot_sum_score`` and the CFD layers, and returns ``{guide index: [off-target, ...]}``."""
        if external_genome:
            index_path = external_genome.bowtie_index
        elif self.genome:
            index_path = self.genome.bowtie_index
        else:
            raise ValueError("No genome / locus specified.")
        if not index_path:
            raise ValueError("Bowtie index is not built for the genome / locus.")

        found = run_bowtie(guides=self.guides, pam=self.pam, genome_index_path=index_path, **kwargs)
        mm_scores, pam_scores = load_cfd_scoring_matrix()
        for ix, G in enumerate(self.guides):
            if ix in found:
                G.off_target_str = get_off_targets_string(found[ix])
                G._attach_off_targets(found[ix], mm_scores, pam_scores, nan_if_empty=True)
        return found

    # ------------------------------------------------------------------ exports (thin)
    def guides_to_dataframe(self):
        return _guides_to_dataframe(self.guides)

    def guides_to_csv(self, filename):
        if not filename:  # reference locus.py:694-700
            raise ValueError("Filename required to save CSV with gRNAs.")
        return _guides_to_csv(self.guides, filename)

    def guides_to_bed(self, filename):
        if not filename:  # reference locus.py:702-707
            raise ValueError("Filename required to save BED with gRNAs.")
        return _guides_to_bed(self.guides, filename)

    # ------------------------------------------------------------------ layers (host-side, outside the hot path)
    def add_layer(self, name, layer_data, layer_pos=None, apply_to_guides=True, is_variation=False):
        """Per-base layer of the locus; clipped to every guide's 23 bases like the reference
        (locus.py:515-557, 409-415).  Variation layers (``is_variation`` with ``layer_pos``) need
        scikit-allel statistics that are outside this package's scope."""
        self._layers[name] = layer_data
        if not apply_to_guides:
            raise ValueError("No guides to apply the data to.")  # (the reference raises on this branch as well)
        if len(self.guides) > 0:
            if is_variation and layer_pos:
                raise NotImplementedError("variation layers (scikit-allel statistics, reference locus.py:433-503) are "
                                          "outside the scope of guido_b200")
            if layer_data.shape[0] != self.length:
                raise ValueError(f"Layer and locus not the same lenght ({layer_data.shape[0]}, {self.length}).")
            for g in self.guides:
                ix = g.guide_start - self.start
                g.add_layer(name, layer_data[ix:ix + 23])

    def rank_guides(self, *args, **kwargs):
        """TOPSIS ranking over guide layers (reference locus.py:622-688) needs the external ``mcdm``
        package and is outside the scope of guido_b200."""
        raise NotImplementedError("rank_guides (mcdm / TOPSIS) is outside the scope of guido_b200")

    def guides_detailed_table(self, filename):
        """The Jinja report of the reference (helpers.py:95-107) is outside the scope of guido_b200."""
        if not filename:
            raise ValueError("Filename required to save CSV with gRNAs.")
        raise NotImplementedError("guides_detailed_table (Jinja template report) is outside the scope of guido_b200")

    def add_azimuth_score(self):
        """Azimuth score of every guide (reference locus.py:718-727); see Guide.add_azimuth_score."""
        return [g.add_azimuth_score() for g in self.guides]


# ---------------------------------------------------------------------- locus factories

def _prepare_annotation(annotation_file_abspath, as_df=True):
    """Annotation table with the reference's renames (GTF gene_id->ID, exon_number->Exon; GFF3
    Name->Exon)."""
    return renamed(read_annotation(annotation_file_abspath), annotation_file_abspath)


def locus_from_coordinates(genome, chromosome, start, end):
    """Locus for ``chromosome:start-end`` (1-based, inclusive); annotated when the genome has a
    GTF/GFF3 (features clipped to the region, pyranges ``intersect`` conventions)."""
    locus_sequence = FastaFile(str(genome.genome_file_abspath)).get_seq(chromosome, start, end)
    ann_path = genome.annotation_file_abspath
    if ann_path and ann_path.exists():
        table = intersect(read_annotation(ann_path), chromosome, start, end)
        table = renamed(table, ann_path)
        if len(table) > 0:
            loc = Locus(genome=genome, sequence=locus_sequence, annotation=table)
            loc._image_region = (chromosome, int(start) - 1)
            return loc
    loc = Locus(genome=genome, sequence=locus_sequence, annotation=None)
    loc._image_region = (chromosome, int(start) - 1)
    return loc


def locus_from_sequence(sequence, sequence_name=None):
    return Locus(sequence=sequence, name=sequence_name)


def locus_from_gene(genome, gene_name):
    """Locus spanning gene ``gene_name`` of the genome's annotation."""
    if not genome.annotation_file_abspath.exists():
        raise ValueError("Annotation file not valid.")
    try:
        ann = _prepare_annotation(genome.annotation_file_abspath)
        gene = ann[(ann["ID"] == gene_name) & ann["Feature"].isin(["protein_coding_gene", "gene"])]
        chromosome = gene.Chromosome.values[0]
        if len(gene) != 1:
            raise TypeError("cannot convert the series to <class 'int'>")  # upstream: int(Series)
        start = int(gene.Start.iloc[0]) + 1  # table starts are 0-based
        end = int(gene.End.iloc[0])
        related = ann["ID"] == gene_name
        for col in ("Parent", "gene_id"):
            if col not in ann.columns:
                raise KeyError(col)  # upstream's query fails on a missing column
            related |= ann[col] == gene_name
        inside = (((ann["Start"] >= start) & (ann["Start"] <= end)) | ((ann["End"] >= start) & (ann["End"] <= end)))
        locus_annotation = ann[related & (ann["Chromosome"] == chromosome) & inside]
        locus_sequence = FastaFile(str(genome.genome_file_abspath)).get_seq(chromosome, start, end)
        loc = Locus(genome=genome, sequence=locus_sequence, annotation=locus_annotation)
        loc._image_region = (chromosome, start - 1)
        return loc
    except Exception as e:
        raise ValueError(f"Gene {gene_name} not found. (Error: {e})")
