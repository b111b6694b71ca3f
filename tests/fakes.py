"""Oracle-backed stand-ins for the three GPU entry points -- TESTS ONLY.

They let the `-m "not gpu"` suite exercise the package's HOST logic (record building, filters,
coordinate conventions) on a box without a GPU.  They are never importable from guido_b200.
"""
import json
import gzip
import os

import numpy as np

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
OT_HIT_DTYPE = np.dtype([("guide", "<u4"), ("gpos", "<u4"), ("hi", "<u4"), ("lo", "<u4")])


def unjson(x):
    if isinstance(x, dict):
        if set(x) == {"__float__"}:
            return float("nan") if x["__float__"] == "nan" else float.fromhex(x["__float__"])
        return {k: unjson(v) for k, v in x.items()}
    if isinstance(x, list):
        return [unjson(v) for v in x]
    return x


def load_golden(name):
    path = os.path.join(GOLDEN, name)
    opener = gzip.open if name.endswith(".gz") else open
    with opener(path, "rt") as fh:
        return unjson(json.load(fh))


def int_keys(d):
    return {int(k): v for k, v in d.items()}


def fake_pam_scan_seq(seq, pam, device=None, mode=0, stats=None):
    return O.pam_scan(seq, pam)


def fake_mmej_batch(windows, cuts, device=None, stats=None):
    tuples, offsets, ncand = [], [0], []
    for w, c in zip(windows, cuts):
        kept, nc = O.mmej_tuples(w, int(c))
        if kept is None:
            kept = np.zeros((0, 4), np.int64)
        packed = (kept[:, 0] | (kept[:, 1] << 8) | (kept[:, 2] << 16) | (kept[:, 3] << 24)).astype(np.uint32)
        tuples.append(packed)
        offsets.append(offsets[-1] + len(packed))
        ncand.append(nc)
    flat = np.concatenate(tuples) if tuples else np.zeros(0, np.uint32)
    return flat, np.array(offsets, np.int64), np.array(ncand, np.int32)


class FakeImage:
    """Answers offtarget_search / locate / contigs like _native.Image, from the oracle."""

    def __init__(self, contigs):
        self._c = contigs
        self.contigs = []
        g = 64
        for name, s in contigs:
            self.contigs.append((name, len(s), g))
            g = ((g + len(s) + 63) // 64) * 64 + 64

    def locate(self, gpos):
        gpos = np.asarray(gpos, dtype=np.int64)
        offs = np.array([c[2] for c in self.contigs], dtype=np.int64)
        ix = np.searchsorted(offs, gpos, side="right") - 1
        return ix, gpos - offs[ix]

    def offtarget_search(self, protos, pam="NGG", core_length=10, core_mismatches=2, total_mismatches=4,
                         algo=0, want_raw_flags=True, stats=None):
        protos = list(protos)
        seed_mms, seed_len, e = O.bowtie_args(pam, core_length, core_mismatches, total_mismatches)
        reads = [O.rev_comp(p + pam) for p in protos]
        raw = O.bowtie_n(self._c, reads, seed_mms, seed_len, e)
        R = 20 + len(pam)
        fixed = {21 + i for i, c in enumerate(pam) if c in "ACGT"}
        names = [c[0] for c in self.contigs]
        seqs = {n: s.upper() for n, s in self._c}
        flags = np.zeros(len(protos), np.uint8)
        rows = []
        for h in raw:
            flags[h["read_ix"]] = 1
            if any((R - off) in fixed for off, _, _ in h["mm"]):
                continue
            w = seqs[h["contig"]][h["offset"]:h["offset"] + R]
            minus = h["strand"] == "+"
            if minus:
                w = O.rev_comp(w)
            hi = lo = 0
            for i, ch in enumerate(w):
                c = "ACGT".index(ch)
                lo |= (c & 1) << i
                hi |= (c >> 1) << i
            if minus:
                hi |= 1 << 31
            rows.append((h["read_ix"], self.contigs[names.index(h["contig"])][2] + h["offset"], hi, lo))
        rows.sort(key=lambda r: (r[0], r[1], r[2] >> 31))
        return np.array(rows, dtype=OT_HIT_DTYPE), (flags if want_raw_flags else None)
