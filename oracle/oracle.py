"""CPU oracle for the guido hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; nothing under ``guido_b200/`` does.

It wraps the C restatement in ``guido_oracle.c`` (built into ``oracle/_ref/`` by
``oracle/Makefile``) and restates, in plain Python, the record-building halves of the
reference so that expected *records* (not only hit sets) exist on a box where
``/root/reference`` is absent.  Every function cites the reference lines it follows
(paths relative to /root/reference).

Pinning status
--------------
* PAM scan, Guide geometry, MMEJ, CFD, bowtie-output parsing/filtering: pinned against the
  reference itself (run in the build container through ``ref_loader``) and against the
  reference's stored outputs -- see ``tests/golden/`` and ``tests/test_oracle_vs_reference.py``.
* The bowtie alignment rule (``bowtie_n``): bowtie 1.3.1 is an external binary that is not in
  this image; the rule restates the bowtie 1 manual.  **Parity unpinned** for that step.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libguido_oracle.so")
_lib = None


def build(force=False):
    """Compile guido_oracle.c -> oracle/_ref/libguido_oracle.so (gcc, seconds)."""
    src = os.path.join(_HERE, "guido_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "_ref/libguido_oracle.so"], check=True,
                   stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _OrcHit(ctypes.Structure):
    _fields_ = [
        ("read_ix", ctypes.c_int32),
        ("contig", ctypes.c_int32),
        ("offset", ctypes.c_int64),
        ("strand", ctypes.c_int32),
        ("n_mm", ctypes.c_int32),
        ("mm_off", ctypes.c_uint8 * 32),
        ("mm_ref", ctypes.c_char * 32),
        ("mm_read", ctypes.c_char * 32),
    ]


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.orc_pam_scan.restype = ctypes.c_int
        L.orc_pam_scan.argtypes = [
            ctypes.c_char_p, ctypes.c_int64, ctypes.c_char_p,
            ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64),
            ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
        L.orc_bowtie_n.restype = ctypes.c_int64
        L.orc_bowtie_n.argtypes = [
            ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_int64), ctypes.c_int,
            ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
        L.orc_mmej.restype = ctypes.c_int64
        L.orc_mmej.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                               ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
        L.orc_threads.restype = ctypes.c_int
        L.orc_set_threads.argtypes = [ctypes.c_int]
        L.orc_seed_index_build.restype = ctypes.c_void_p
        L.orc_seed_index_build.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_int64),
                                           ctypes.POINTER(ctypes.c_int64), ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
        L.orc_seed_index_free.argtypes = [ctypes.c_void_p]
        L.orc_seed_index_size.restype = ctypes.c_uint64
        L.orc_seed_index_size.argtypes = [ctypes.c_void_p]
        L.orc_seed_index_search.restype = ctypes.c_int64
        L.orc_seed_index_search.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_int64]
        _lib = L
    return _lib


def threads():
    return int(lib().orc_threads())


def set_threads(n):
    """Number of OpenMP threads of the C oracle (torchrun exports OMP_NUM_THREADS=1: the CPU arm of
    bench.py sets the count it wants explicitly)."""
    lib().orc_set_threads(int(n))
    return threads()


class SeedIndex:
    """CPU seed-indexed off-target search (throughput baseline of bench.py; see guido_oracle.c).

    ``contigs``: list of (name, sequence bytes/str, global offset).  Building is the counterpart
    of ``bowtie-build`` (guido/genome.py:157-164); ``search`` the counterpart of the bowtie call of
    guido/off_targets.py:151-167 followed by guido's PAM filter (:169-198)."""

    def __init__(self, contigs, pam="NGG", core_length=10):
        seqs = [c[1].encode() if isinstance(c[1], str) else bytes(c[1]) for c in contigs]
        self._keep = seqs
        arr = (ctypes.c_char_p * len(seqs))(*seqs)
        lens = (ctypes.c_int64 * len(seqs))(*[len(s) for s in seqs])
        offs = (ctypes.c_int64 * len(seqs))(*[int(c[2]) for c in contigs])
        self._h = lib().orc_seed_index_build(arr, lens, offs, len(seqs), pam.encode(), core_length)
        if not self._h:
            raise ValueError("seed index: bad PAM / core length, or more than 2^32 windows")
        self.n_windows = int(lib().orc_seed_index_size(self._h))
        self._keep = None  # the index holds what it needs

    def search(self, protos_u8, core_mismatches=2, total_max=4, want_hits=False):
        """protos_u8: (n, 20) uint8 letters.  Returns the hit count, or a sorted uint64 array
        guide << 33 | global position << 1 | strand when ``want_hits``."""
        blob = np.ascontiguousarray(protos_u8, dtype=np.uint8)
        n = blob.shape[0]
        if not want_hits:
            return int(lib().orc_seed_index_search(self._h, blob.tobytes(), n, core_mismatches, total_max, None, 0))
        cap = max(1024, 64 * n)
        while True:
            out = np.empty(cap, np.uint64)
            got = int(lib().orc_seed_index_search(self._h, blob.tobytes(), n, core_mismatches, total_max,
                                                  out.ctypes.data, cap))
            if got <= cap:
                return np.sort(out[:got])
            cap = got

    def close(self):
        if self._h:
            lib().orc_seed_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------------- helpers

_COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def rev_comp(seq):
    """guido/helpers.py:19-26 (rna=False): anything outside ACGTN becomes N."""
    return "".join(_COMP.get(b, "N") for b in reversed(seq))


# --------------------------------------------------------------------------- PAM scan

def pam_scan(seq, pam):
    """guido/locus.py:178-183 -> (pams_fw, pams_rv) as int32 arrays of 0-based starts."""
    b = seq.encode() if isinstance(seq, str) else bytes(seq)
    n = len(b)
    cap = max(n, 1)
    fw = np.empty(cap, np.int32)
    rv = np.empty(cap, np.int32)
    nf, nr = ctypes.c_int64(0), ctypes.c_int64(0)
    rc = lib().orc_pam_scan(b, n, pam.encode(), fw.ctypes.data, cap, ctypes.byref(nf),
                            rv.ctypes.data, cap, ctypes.byref(nr))
    if rc == -1:
        raise KeyError(pam)  # the reference's iupac_dict lookup fails the same way
    assert rc == 0
    return fw[:nf.value].copy(), rv[:nr.value].copy()


def guide_records(seq, start, pam, chromosome="seq", max_flank=75, cut_offset=3):
    """guido/locus.py:158-211 + guido/guides.py:17-106 as plain dict records.

    Returns what ``Locus._find_guides_in_interval(seq, start, pam)`` returns (fw guides by
    ascending PAM position, then rv guides; guides whose guide_seq holds an 'N' dropped),
    each as a dict of the Guide attributes the hot path defines.
    """
    S = seq.upper()
    P = len(pam)
    fw, rv = pam_scan(S, pam)
    out = []
    for strand, positions in (("+", fw), ("-", rv)):
        for p in positions.tolist():
            if strand == "+":
                cut = p - cut_offset
                rs, re_ = cut - 17, cut + cut_offset + P
                gseq = S[slice(rs, re_)]
                glong = S[slice(rs - 4, re_ + 3)]
            else:
                cut = p + P + cut_offset
                rs, re_ = p, cut + 17
                gseq = rev_comp(S[slice(rs, re_)])
                glong = rev_comp(S[slice(rs - 3, re_ + 4)])
            left, right, rel = cut - max_flank, cut + max_flank, max_flank
            if left < 0:
                left, rel = 0, cut
            if right > len(S):
                right = len(S)
            if "N" in gseq:
                continue
            out.append({
                "guide_seq": gseq, "guide_long_seq": glong, "guide_pam": gseq[-P:],
                "guide_chrom": chromosome, "guide_start": start + rs, "guide_end": start + re_ - 1,
                "guide_strand": strand, "relative_cut_pos": rel, "absolute_cut_pos": start + cut,
                "locus_seq": S[slice(left, right)],
            })
    return out


def find_guides_records(seq, start=1, end=None, pam="NGG", min_flanking_length=0,
                        chromosome=None, intervals=None):
    """guido/locus.py:275-329 (the 'all' features path, or explicit ``intervals``)."""
    if end is None:
        end = start + len(seq)  # guido/locus.py:94
    if intervals is None:
        intervals = [[start, end]]
    guides = []
    for istart, iend in intervals:
        a = max(istart - min_flanking_length - start, 0)
        b = min(iend + min_flanking_length - start, len(seq))
        for g in guide_records(seq[a:b], istart - min_flanking_length, pam, chromosome):
            if len(g["guide_seq"]) == 23 and istart <= g["absolute_cut_pos"] <= iend:
                guides.append(g)
    guides = sorted(guides, key=lambda g: g["absolute_cut_pos"])
    for i, g in enumerate(guides):
        g["id"] = f"gRNA-{i + 1}"
    return guides


# --------------------------------------------------------------------------- off-targets

def bowtie_n(contigs, reads, seed_mms, seed_len, qual_ceiling):
    """Every alignment bowtie 1 would report for `-n seed_mms -l seed_len -e qual_ceiling -a -y`.

    contigs: list of (name, sequence str); reads: list of equal-length str (A/C/G/T/N).
    Returns a list of dicts: read_ix, strand ('+'/'-' as bowtie prints), contig (name), offset,
    mm = [(offset_from_read_5prime, ref_base_fwd, read_base_fwd), ...].
    """
    if not reads:
        return []
    R = len(reads[0])
    assert all(len(r) == R for r in reads)
    seqs = [s.encode() for _, s in contigs]
    arr = (ctypes.c_char_p * len(seqs))(*seqs)
    lens = (ctypes.c_int64 * len(seqs))(*[len(s) for s in seqs])
    blob = "".join(reads).encode()
    cap = 1 << 16
    while True:
        buf = (_OrcHit * cap)()
        n = lib().orc_bowtie_n(arr, lens, len(seqs), blob, len(reads), R, seed_mms, seed_len,
                               qual_ceiling, ctypes.addressof(buf), cap)
        assert n >= 0
        if n <= cap:
            break
        cap = int(n)
    out = []
    for i in range(n):
        h = buf[i]
        ref = bytes(bytearray(ctypes.string_at(ctypes.addressof(h) + _OrcHit.mm_ref.offset, 32)))
        rd = bytes(bytearray(ctypes.string_at(ctypes.addressof(h) + _OrcHit.mm_read.offset, 32)))
        out.append({
            "read_ix": int(h.read_ix), "strand": chr(h.strand), "contig": contigs[h.contig][0],
            "offset": int(h.offset),
            "mm": [(int(h.mm_off[k]), chr(ref[k]), chr(rd[k])) for k in range(h.n_mm)],
        })
    return out


def bowtie_n_count(contigs, reads, seed_mms, seed_len, qual_ceiling):
    """Same search, only the number of alignments (for timing the compare loop)."""
    R = len(reads[0])
    seqs = [s.encode() if isinstance(s, str) else s for _, s in contigs]
    arr = (ctypes.c_char_p * len(seqs))(*seqs)
    lens = (ctypes.c_int64 * len(seqs))(*[len(s) for s in seqs])
    blob = "".join(reads).encode()
    return int(lib().orc_bowtie_n(arr, lens, len(seqs), blob, len(reads), R, seed_mms, seed_len,
                                  qual_ceiling, None, 0))


def bowtie_tsv(raw_hits, read_names, reads):
    """Format alignments the way bowtie 1 prints them (8 tab-separated columns), the shape
    guido/off_targets.py:174-181 splits.  Validated against test.ipynb:14079-14082."""
    lines = []
    for h in raw_hits:
        rd = reads[h["read_ix"]]
        shown = rd if h["strand"] == "+" else rev_comp(rd)
        desc = ",".join(f"{o}:{r}>{q}" for o, r, q in h["mm"])
        lines.append("\t".join([read_names[h["read_ix"]], h["strand"], h["contig"],
                                str(h["offset"]), shown, "I" * len(rd), "0", desc]))
    return "\n".join(lines) + ("\n" if lines else "")


def reads_for_guides(guide_seqs, pam):
    """guido/off_targets.py:128-139: read_i = rev_comp(protospacer + PAM pattern)."""
    P = len(pam)
    return [rev_comp(g[:-P] + pam) for g in guide_seqs]


def bowtie_args(pam, core_length, core_mismatches, total_mismatches):
    """guido/off_targets.py:128,141-158 -> (seed_mms, seed_len, qual_ceiling); same ValueErrors."""
    pam_mismatches = rev_comp(pam).count("N")
    if core_mismatches + pam_mismatches > 3:
        raise ValueError(
            f"The value for the parameter core_mismatches is not valid: {core_mismatches}")
    if core_mismatches > total_mismatches:
        raise ValueError(
            "The value for core_mismatches cannot be greater than total_mismatches: "
            f"{core_mismatches} > {total_mismatches}")
    return core_mismatches + pam_mismatches, core_length + len(pam), total_mismatches * 30 + 30


def off_target_records(guides, contigs, pam="NGG", core_length=10, core_mismatches=2,
                       total_mismatches=4):
    """guido/off_targets.py:89-227 with the bowtie subprocess replaced by ``bowtie_n``.

    guides: list of dicts with guide_seq, guide_chrom, guide_start (as find_guides_records gives).
    Returns {ix: [record, ...]} -- key present for any raw alignment (off_targets.py:183-184).
    Records are sorted canonically (chromosome, start, strand) because bowtie's emission order is
    unspecified.
    """
    seed_mms, seed_len, e = bowtie_args(pam, core_length, core_mismatches, total_mismatches)
    reads = reads_for_guides([g["guide_seq"] for g in guides], pam)
    raw = bowtie_n(contigs, reads, seed_mms, seed_len, e)
    fixed = [21 + i for i, c in enumerate(pam) if c in "ATGC"]  # off_targets.py:169-171
    out = {}
    for h in raw:
        ix = h["read_ix"]
        g = guides[ix]
        out.setdefault(ix, [])
        R = len(reads[ix])
        if h["strand"] == "-":  # off_targets.py:186-192
            strand, t_seq = "+", rev_comp(reads[ix])
        else:
            strand, t_seq = "-", rev_comp(reads[ix])
        # NB: bowtie prints the read reverse-complemented for '-' rows; the reference then
        # rev_comps the '+' rows, so t_seq is the read in guide orientation either way.
        if not h["mm"]:
            # bowtie prints an empty descriptor field for a perfect alignment (only possible when
            # the PAM pattern has no N); _parse_mismatches("") then fails at off_targets.py:73
            raise ValueError("not enough values to unpack (expected 2, got 1)")
        mm = {}
        for off, ref, _ in h["mm"]:  # off_targets.py:64-80
            mm[R - off] = rev_comp(ref) if strand == "-" else ref
        if not mm or any(p in fixed for p in mm):
            continue
        if not (str(g["guide_chrom"]) != h["contig"] and int(g["guide_start"]) != h["offset"] + 1):
            continue  # off_targets.py:199
        out[ix].append({
            "ix": ix, "mismatches": mm,
            "mismatches_string": "".join(mm.get(i + 1, ".") for i in range(R)),
            "chromosome": h["contig"], "start": h["offset"], "strand": strand,
            "seq": "".join(mm.get(i + 1, t_seq[i]) for i in range(R)),
        })
    for ix in out:
        out[ix].sort(key=lambda r: (r["chromosome"], r["start"], r["strand"]))
    return out


def cfd_score(guide_seq, guide_pam, ot_seq, mm_scores, pam_scores):
    """guido/off_targets.py:31-61 for one off-target (float64, sequential product)."""
    wt = guide_seq[:-len(guide_pam)].replace("T", "U")
    sg = ot_seq[:20].replace("T", "U")
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "U": "A", "-": "-"}
    score = 1
    for i, sl in enumerate(sg):
        if wt[i] != sl:
            score *= mm_scores["r" + wt[i] + ":d" + comp.get(sl, "N") + "," + str(i + 1)]
    score *= pam_scores[guide_pam[1:]]
    return score


# --------------------------------------------------------------------------- MMEJ

def mmej_tuples(seq, cut):
    """guido/mmej.py:8-64 -> (kept (n,4) int array in reference order | None, n_candidates)."""
    b = seq.encode()
    cap = 1 << 14
    while True:
        out = np.empty((cap, 4), np.uint16)
        nc = ctypes.c_int64(0)
        n = lib().orc_mmej(b, len(b), cut, out.ctypes.data, cap, ctypes.byref(nc))
        if n < 0:
            return None, 0
        if n <= cap:
            return out[:n].astype(np.int64), int(nc.value)
        cap = int(n)


def mmej_patterns(rel_cut, seq, length_weight):
    """guido/mmej.py:8-120: full list of pattern dicts (or None)."""
    kept, _ = mmej_tuples(seq, rel_cut)
    if kept is None:
        return None
    left, right = seq[:rel_cut], seq[rel_cut:]
    res = []
    for ls, le, rs, re_ in kept:  # numpy int64 like the reference's combinations_array rows
        pat = left[ls:le]
        gc = pat.count("G") + pat.count("C")
        ldel = len(left) - ls
        dlen = ldel + rs
        dseq = left[ls:] + right[:rs]
        lf = round(1 / math.exp(dlen / length_weight), 3)
        score = 100 * lf * ((len(pat) - gc) + (gc * 2))
        res.append({
            "left": left[:ls] + "-" * ldel, "left_seq": left[:ls], "left_seq_position": ls,
            "right": "+" * rs + right[rs:], "right_seq": right[rs:],
            "right_seq_position": rs + len(dseq), "pattern": pat, "pattern_len": len(pat),
            "pattern_score": score, "deletion_seq": dseq,
            "frame_shift": "-" if dlen % 3 == 0 else "+",
        })
    return res


def simulate_end_joining(locus_seq, rel_cut, n_patterns=5, length_weight=20):
    """guido/guides.py:132-182 -> (patterns | None, mmej_str | None, sum_score, oof_score).

    Raises TypeError where the reference does (no pattern at all, guides.py:173)."""
    if "N" in locus_seq:
        return None, None, 0, 0
    pats = mmej_patterns(rel_cut, locus_seq, length_weight)
    if not pats:
        raise TypeError("'NoneType' object is not iterable")
    top = sorted(pats, key=lambda x: -x["pattern_score"])[:n_patterns]
    oof_sum = sum([p["pattern_score"] for p in top if p["frame_shift"] == "+"])
    sum_score = sum([p["pattern_score"] for p in top])
    return top, "|".join(p["frame_shift"] for p in top), sum_score, oof_sum / sum_score * 100
