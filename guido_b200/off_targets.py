"""Off-target search and scoring (drop-in for guido.off_targets, reference guido/off_targets.py).

``run_bowtie`` keeps its name, signature and return format, but no bowtie runs: the guides are
searched against the genome image resident in HBM by ``guido_offtarget_search`` (C-ABI), which
reports the alignments that bowtie 1 ``-n/-l/-e -a -y`` reports ACCORDING TO THE BOWTIE 1 MANUAL
(seed = the first ``-l`` read bases with at most ``-n`` mismatches, 30 per mismatch against ``-e``,
a read N mismatches everything, no alignment over ambiguous reference bases) and that guido keeps
(off_targets.py:151-158,169-199).  bowtie 1.3.1 itself is not available where this package was
built, so that reading of the rule is pinned against the oracle only ("parity unpinned",
DESIGN.md section 2); ``tools/compare_with_bowtie.py`` diffs it against a real bowtie where one is
installed.  ``genome_index_path`` is the image path that ``Genome.build``
stores in ``Genome.bowtie_index``; ``threads`` and ``bowtie_path`` are accepted and ignored.
"""
import os

import numpy as np

from . import _native
from .helpers import rev_comp

_LETTERS = np.frombuffer(b"ACGT", dtype=np.uint8)
_image_cache = {}


def get_image(index_path, device=None, shard_ix=0, n_shards=1):
    """Load (once) and upload the genome image stored at ``index_path``."""
    path = os.fspath(index_path)
    dev = _native.default_device() if device is None else device
    key = (os.path.abspath(path), dev, shard_ix, n_shards)
    mtime = os.path.getmtime(path)
    hit = _image_cache.get(key)
    if hit is None or hit[0] != mtime:
        img = _native.Image.load(path).upload(device=dev, shard_ix=shard_ix, n_shards=n_shards)
        _image_cache[key] = (mtime, img)
    return _image_cache[key][1]


def register_image(index_path, image):
    """Make an already uploaded image answer for ``index_path`` (used by Genome.build and tests)."""
    path = os.fspath(index_path)
    inf = image.info
    mtime = os.path.getmtime(path) if os.path.exists(path) else 0.0
    _image_cache[(os.path.abspath(path), inf.device, 0, 1)] = (mtime, image)


def calculate_ot_sum_score(off_targets, ot_mismatch_weights=[10, 5, 4, 3, 1], pam="NGG"):
    shift = pam.count("N")
    return sum([ot_mismatch_weights[len(ot["mismatches"]) - shift] for ot in off_targets])


def get_off_targets_string(off_targets, pam="NGG"):
    counts = [0, 0, 0, 0, 0]
    shift = pam.count("N")
    for ot in off_targets:
        counts[len(ot["mismatches"]) - shift] += 1
    return "|".join(str(c) for c in counts)


def calculate_cfd_score(guide, offtargets, mm_scores, pam_scores):
    """CFD score (Doench et al. 2016) of each off-target: sequential float64 product over the 20
    protospacer positions, times the score of the GUIDE's own PAM dinucleotide
    (reference off_targets.py:31-61)."""
    wt = guide.guide_seq[: -len(guide.guide_pam)].replace("T", "U")
    pam_factor_key = guide.guide_pam[1:]
    scores = []
    for ot in offtargets:
        sg = ot["seq"][:20].replace("T", "U")
        score = 1
        for i, base in enumerate(sg):
            if wt[i] != base:
                score *= mm_scores["r" + wt[i] + ":d" + rev_comp(base, rna=True) + "," + str(i + 1)]
        score *= pam_scores[pam_factor_key]
        scores.append(score)
    return np.array(scores)


def decode_hits(hits, width):
    """hits (structured array) -> (n, width) uint8 letters of the genomic window, guide orientation."""
    shifts = np.arange(width, dtype=np.uint32)
    lo = (hits["lo"][:, None] >> shifts) & 1
    hi = (hits["hi"][:, None] >> shifts) & 1
    return _LETTERS[(lo | (hi << 1)).astype(np.intp)]


def records_from_hits(guides, hits, raw_flags, image, pam):
    """C-ABI hits -> the reference's ``{ix: [off_target dict, ...]}`` (off_targets.py:174-225)."""
    P = len(pam)
    width = 20 + P
    out = {}
    if raw_flags is not None:
        for ix in np.nonzero(raw_flags)[0].tolist():
            out[ix] = []
    if len(hits) == 0:
        return out
    letters = decode_hits(hits, width)
    contig_ix, offset = image.locate(hits["gpos"])
    names = [c[0] for c in image.contigs]
    strand_minus = (hits["hi"] >> 31).astype(bool)
    ambiguous_pam = [k for k, c in enumerate(pam) if c not in "ACGT"]
    proto = np.frombuffer("".join(g.guide_seq[:20] for g in guides).encode(), dtype=np.uint8).reshape(-1, 20)
    differs = letters[:, :20] != proto[hits["guide"]]
    if not ambiguous_pam and not differs.any(axis=1).all():
        # Upstream quirk, reproduced: with a PAM pattern without N a perfect alignment reaches
        # _parse_mismatches("") and run_bowtie dies there (off_targets.py:73).
        raise ValueError("not enough values to unpack (expected 2, got 1)")
    # every guide that owns a hit has a key (off_targets.py:183-184) ...
    for ix in np.unique(hits["guide"]).tolist():
        out.setdefault(int(ix), [])
    # ... but guido keeps an alignment only on another chromosome (off_targets.py:199), and only if it
    # carries at least one mismatch entry: both tests are done for all hits at once, the Python loop
    # below sees the survivors only (a locus' hits are mostly the guides' own sites)
    name_ix = {n: i for i, n in enumerate(names)}
    g_chrom = np.array([name_ix.get(str(G.guide_chrom), -1) for G in guides], dtype=np.int64)
    g_start = np.array([int(G.guide_start) for G in guides], dtype=np.int64)
    hg = hits["guide"].astype(np.int64)
    keep = (g_chrom[hg] != contig_ix) & (g_start[hg] != offset + 1)
    if not ambiguous_pam:
        keep &= differs.any(axis=1)
    for h in np.nonzero(keep)[0].tolist():
        ix = int(hits["guide"][h])
        chrom = names[int(contig_ix[h])]
        start = int(offset[h])
        window = letters[h].tobytes().decode()
        mismatches = {}
        for k in reversed(ambiguous_pam):       # descending position = ascending bowtie offset
            mismatches[21 + k] = window[20 + k]
        for i in np.nonzero(differs[h])[0][::-1].tolist():
            mismatches[i + 1] = window[i]
        if not mismatches:
            continue
        out[ix].append({
            "ix": ix,
            "mismatches": mismatches,
            "mismatches_string": "".join(window[i] if (i + 1) in mismatches else "." for i in range(width)),
            "chromosome": chrom,
            "start": start,
            "strand": "-" if strand_minus[h] else "+",
            "seq": window,
        })
    return out


def run_bowtie(guides, pam="NGG", core_length=10, core_mismatches=2, total_mismatches=4,
               genome_index_path=None, threads=1, bowtie_path="bin/bowtie/", algo=_native.OT_AUTO):
    """Find off-targets for a list of guides.

    Same parameters as the reference: ``core_length`` PAM-proximal bases may hold at most
    ``core_mismatches`` mismatches, the protospacer at most ``total_mismatches``; PAM letters
    A/C/G/T must match exactly.  Returns ``{guide index: [off-target dict, ...]}``; a guide
    without any raw alignment in the genome has no key.
    """
    pam_length = len(pam)
    protos = []
    for G in guides:
        p = G.guide_seq[:-pam_length]
        if len(p) != 20:
            raise ValueError(f"only 20-nt protospacers are supported, got {len(p)} nt: {G.guide_seq!r}")
        protos.append(p)
    # argument checks and messages of off_targets.py:141-149 (raised by the C-ABI as GUIDO_ERR_ARG)
    image = get_image(genome_index_path)
    hits, raw_flags = image.offtarget_search(protos, pam=pam, core_length=core_length,
                                             core_mismatches=core_mismatches,
                                             total_mismatches=total_mismatches, algo=algo)
    return records_from_hits(guides, hits, raw_flags, image, pam)
