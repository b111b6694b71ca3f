"""``Guide`` -- the gRNA record (drop-in for guido.guides.Guide, reference guido/guides.py).

The record format is the reference's: same constructor signature, same attribute names, same
coordinate conventions (1-based inclusive guide_start/guide_end, guides.py:69-90), same quirks
(Python slice wrap-around for PAMs too close to the interval start, ``__getattr__`` raising
``TypeError``).  What changes is where the heavy lifting happens: MMEJ enumeration and the
off-target search run on the GPU through the C-ABI (``_native``); this class only assembles records.
"""
from collections import Counter

import numpy as np

from . import mmej as _mmej
from . import off_targets as _ot
from .helpers import load_cfd_scoring_matrix, rev_comp


def guide_geometry(sequence, pam_position, pam_len, strand, max_flanking_length=75, cut_offset=3, start=0,
                   rc_sequence=None):
    """All position-derived fields of a guide (reference guido/guides.py:50-93).

    ``rc_sequence`` (optional) is the reverse complement of the whole ``sequence``; with it the
    reverse-strand strings are plain slices instead of per-guide reverse complements.
    """
    n = len(sequence)
    if strand == "-":
        cut = pam_position + pam_len + cut_offset
        rel_start, rel_end = pam_position, cut + 17
        if rc_sequence is not None and rel_start - 3 >= 0:
            # rc(S[a:b]) == RC[n-b:n-a] whenever 0 <= a <= b (b may exceed n: both clamp alike)
            guide_seq = rc_sequence[max(n - rel_end, 0):n - rel_start]
            long_seq = rc_sequence[max(n - (rel_end + 4), 0):n - (rel_start - 3)]
        else:
            guide_seq = rev_comp(sequence[rel_start:rel_end])
            long_seq = rev_comp(sequence[rel_start - 3:rel_end + 4])
    else:
        cut = pam_position - cut_offset
        rel_start, rel_end = cut - 17, cut + cut_offset + pam_len
        guide_seq = sequence[rel_start:rel_end]
        long_seq = sequence[rel_start - 4:rel_end + 3]
    lo, hi, rel_cut = cut - max_flanking_length, cut + max_flanking_length, max_flanking_length
    if lo < 0:
        lo, rel_cut = 0, cut
    if hi > n:
        hi = n
    return {
        "locus_seq": sequence[lo:hi],
        "guide_seq": guide_seq,
        "guide_long_seq": long_seq,
        "guide_pam": guide_seq[-pam_len:],
        "guide_start": start + rel_start,
        "guide_end": start + rel_end - 1,
        "relative_cut_pos": rel_cut,
        "absolute_cut_pos": start + cut,
    }


class Guide:
    def __init__(self, sequence, pam_position, pam_len, strand="+", max_flanking_length=75, cut_offset=3,
                 chromosome="seq", start=0):
        """A gRNA found at ``pam_position`` of ``sequence`` (same parameters as the reference:
        sequence, PAM start in it, PAM length, strand, MMEJ flank, cut offset, chromosome name and
        the coordinate of ``sequence[0]``)."""
        geo = guide_geometry(sequence, pam_position, pam_len, strand, max_flanking_length, cut_offset, start)
        self._fill(geo, chromosome, strand)

    def _fill(self, geo, chromosome, strand):
        d = self.__dict__
        d["locus_seq"] = geo["locus_seq"]
        d["guide_seq"] = geo["guide_seq"]
        d["guide_long_seq"] = geo["guide_long_seq"]
        d["guide_pam"] = geo["guide_pam"]
        d["guide_chrom"] = chromosome
        d["guide_start"] = geo["guide_start"]
        d["guide_end"] = geo["guide_end"]
        d["guide_strand"] = strand
        d["relative_cut_pos"] = geo["relative_cut_pos"]
        d["absolute_cut_pos"] = geo["absolute_cut_pos"]
        d["rank_score"] = 0.0
        d["rank"] = 0
        d["id"] = ""
        d["mmej_patterns"] = []
        d["off_targets"] = []
        d["off_target_str"] = ""
        d["_layers"] = {}

    @classmethod
    def _from_geometry(cls, geo, chromosome, strand):
        g = cls.__new__(cls)
        g._fill(geo, chromosome, strand)
        return g

    @classmethod
    def _from_fields(cls, locus_seq, guide_seq, long_seq, pam_len, chromosome, guide_start, guide_end,
                     strand, rel_cut, abs_cut):
        """Same record as ``_fill`` from already-sliced strings (the bulk path of Locus.find_guides)."""
        g = cls.__new__(cls)
        g.__dict__.update(
            locus_seq=locus_seq, guide_seq=guide_seq, guide_long_seq=long_seq,
            guide_pam=guide_seq[-pam_len:], guide_chrom=chromosome, guide_start=guide_start,
            guide_end=guide_end, guide_strand=strand, relative_cut_pos=rel_cut, absolute_cut_pos=abs_cut,
            rank_score=0.0, rank=0, id="", mmej_patterns=[], off_targets=[], off_target_str="", _layers={})
        return g

    def __repr__(self):
        return (f"{self.id if self.id else 'gRNA'}({self.guide_seq}|{self.guide_chrom}:"
                f"{self.guide_start}-{self.guide_end}|{self.guide_strand}|)")

    def __getattr__(self, attr):
        # the reference's `return self[attr]` (guides.py:115-116): a missing attribute is a TypeError
        raise TypeError(f"'{type(self).__name__}' object is not subscriptable")

    @property
    def location(self):
        return f"{self.guide_chrom}:{self.guide_start}-{self.guide_end}"

    # ------------------------------------------------------------------ MMEJ
    def _create_mmej_oof_string(self, mmej_patterns):
        return "|".join([p["frame_shift"] for p in mmej_patterns])

    def _finish_end_joining(self, patterns, n_patterns):
        """guides.py:150-176 given the full pattern list of this guide's window (or None)."""
        if patterns:
            top = sorted(patterns, key=lambda p: -p["pattern_score"])[:n_patterns]
            oof_sum = sum([p["pattern_score"] for p in top if p["frame_shift"] == "+"])
            sum_score = sum([p["pattern_score"] for p in top])
            oof_score = oof_sum / sum_score * 100
        else:
            top, sum_score, oof_score = None, None, None
        self.mmej_patterns = top
        self.mmej_str = self._create_mmej_oof_string(top)  # TypeError when top is None, as upstream
        self.add_layer("mmej_sum_score", sum_score)
        self.add_layer("mmej_oof_score", oof_score)

    def _skip_end_joining(self):
        self.mmej_patterns = None
        self.mmej_str = None
        self.add_layer("mmej_sum_score", 0)
        self.add_layer("mmej_oof_score", 0)

    def simulate_end_joining(self, n_patterns=5, length_weight=20):
        """Predict MMEJ deletion patterns around the cut site (Bae et al. 2014 scoring) and keep the
        ``n_patterns`` best; adds the ``mmej_sum_score`` / ``mmej_oof_score`` layers."""
        if "N" in self.locus_seq:
            self._skip_end_joining()
            return
        patterns = _mmej.top_patterns_batch([self.relative_cut_pos], [self.locus_seq], length_weight, n_patterns)[0]
        self._finish_end_joining(patterns, n_patterns)

    # ------------------------------------------------------------------ off-targets
    def _attach_off_targets(self, off_targets, mm_scores, pam_scores, nan_if_empty):
        self.off_targets = off_targets
        self.add_layer("ot_sum_score", _ot.calculate_ot_sum_score(off_targets))
        cfd = _ot.calculate_cfd_score(self, off_targets, mm_scores, pam_scores)
        for ot, score in zip(off_targets, cfd.tolist()):
            ot["cfd_score"] = score
        if nan_if_empty and len(cfd) == 0:
            for key in ("ot_cfd_score_mean", "ot_cfd_score_max", "ot_cfd_score_sum"):
                self.add_layer(key, np.nan)
        else:
            self.add_layer("ot_cfd_score_mean", cfd.mean())
            self.add_layer("ot_cfd_score_max", cfd.max())
            self.add_layer("ot_cfd_score_sum", cfd.sum())

    def find_off_targets(self, genome, **kwargs):
        """Search ``genome`` (built, i.e. with a device genome image) for off-targets of this guide;
        fills ``off_targets`` and the ot_sum_score / ot_cfd_score_{mean,max,sum} layers."""
        if not genome:
            raise ValueError("No genome / locus specified.")
        index_path = genome.bowtie_index
        if not index_path:
            raise ValueError("Bowtie index is not built for the genome / locus.")
        found = _ot.run_bowtie(guides=[self], genome_index_path=index_path, **kwargs)
        mm_scores, pam_scores = load_cfd_scoring_matrix()
        self._attach_off_targets(found[0], mm_scores, pam_scores, nan_if_empty=False)

    @property
    def off_targets_string(self):
        """"n0|n1|...|nk (total)": off-target counts by number of mismatches (guides.py:236-273)."""
        if not self.off_targets:
            return "0"
        shift = self.guide_pam.count("N")
        counts = Counter(len(ot["mismatches"].keys()) - shift for ot in self.off_targets)
        body = "|".join(str(counts[i]) for i in range(max(counts.keys()) + 1))
        return f"{body} ({len(self.off_targets)})"

    # ------------------------------------------------------------------ layers
    @property
    def layers(self):
        return self._layers

    def add_layer(self, name, layer_data):
        self._layers[name] = layer_data

    def layer(self, key):
        if key in self._layers.keys():
            return self._layers[key]
        raise ValueError(f"Layer {key} was not added to the gRNA.")

    def add_azimuth_score(self, model_file="V3_model_nopos.pickle"):
        """Azimuth (Doench 2016) on-target score.  Out of scope for the accelerated path: it needs the
        external ``azimuth`` package and its pickled models; without them only the reference's
        "not scorable" branch (0.0) is reproduced."""
        scorable = (len(self.guide_long_seq) == 30 and "N" not in self.guide_long_seq
                    and self.guide_pam.endswith("GG") and len(self.guide_pam) == 3)
        if scorable:
            try:
                from azimuth import model_comparison  # noqa: F401
            except ImportError as exc:
                raise RuntimeError("azimuth is not installed; add_azimuth_score is outside the "
                                   "B200 hot path") from exc
            import os
            score = float(model_comparison.predict(
                np.array([self.guide_long_seq]), length_audit=True, pam_audit=True,
                model_file=os.path.join(os.path.dirname(os.path.realpath(__file__)), "data", model_file))[0])
        else:
            score = 0.0
        self.add_layer("azimuth_score", score)
        return score
