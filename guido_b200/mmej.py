"""MMEJ deletion patterns (drop-in for guido.mmej, reference guido/mmej.py:8-120).

The candidate enumeration and nested-pattern elimination run on the GPU (one warp per cut site,
``guido_mmej_batch``); this module turns the kept ``(left_start, left_end, right_start, right_end)``
tuples into the reference's pattern dicts.  The score is a float64 function of three small
integers and is evaluated here with the reference's exact expression and operation order
(mmej.py:86-91), so scores are bit-identical to the reference (e.g. 230.79999999999998).
"""
import math

import numpy as np

from . import _native


def _pattern_dict(left_seq, right_seq, ls, le, rs, length_weight):
    pattern = left_seq[ls:le]
    gc = pattern.count("G") + pattern.count("C")
    left_del = len(left_seq) - ls
    del_len = left_del + rs
    deletion = left_seq[ls:] + right_seq[:rs]
    length_factor = round(1 / math.exp(del_len / length_weight), 3)
    score = 100 * length_factor * ((len(pattern) - gc) + (gc * 2))
    return {
        "left": left_seq[:ls] + "-" * left_del,
        "left_seq": left_seq[:ls],
        "left_seq_position": ls,
        "right": "+" * rs + right_seq[rs:],
        "right_seq": right_seq[rs:],
        "right_seq_position": rs + len(deletion),
        "pattern": pattern,
        "pattern_len": len(pattern),
        "pattern_score": score,
        "deletion_seq": deletion,
        "frame_shift": "-" if del_len % 3 == 0 else "+",
    }


def _unpack(tuples):
    t = np.asarray(tuples, dtype=np.uint32)
    # np.int64 positions, like the rows of the reference's combinations_array
    return (t & 255).astype(np.int64), ((t >> 8) & 255).astype(np.int64), ((t >> 16) & 255).astype(np.int64)


def patterns_from_tuples(rel_cut_position, sequence, tuples, length_weight):
    left_seq, right_seq = sequence[:rel_cut_position], sequence[rel_cut_position:]
    ls, le, rs = _unpack(tuples)
    return [_pattern_dict(left_seq, right_seq, a, b, c, length_weight) for a, b, c in zip(ls, le, rs)]


def _split_points(rel_cut_positions, sequences):
    """The index at which ``sequence[:cut]`` / ``sequence[cut:]`` split the sequence, with Python's
    slice semantics (a negative cut counts from the end, an out-of-range one clamps) -- what the
    reference's two slices do (mmej.py:13-14).  Shared by both batch entry points.  Each side of the
    split may hold at most 255 letters (the kernel's position bytes); longer flanks raise ValueError
    in the C layer, where the reference would carry on."""
    return [len(seq[:int(c)]) for seq, c in zip(sequences, rel_cut_positions)]


def generate_mmej_patterns_batch(rel_cut_positions, sequences, length_weight, device=None, stats=None):
    """One GPU call for many cut sites; returns a list with, per site, the reference's result of
    ``generate_mmej_patterns`` (a list of pattern dicts, or None when there is no candidate)."""
    sequences = list(sequences)
    cuts = _split_points(rel_cut_positions, sequences)
    if not sequences:
        return []
    tuples, offsets, n_cand = _native.mmej_batch(sequences, cuts, device=device, stats=stats)
    out = []
    for g, (seq, cut) in enumerate(zip(sequences, cuts)):
        if n_cand[g] == 0:
            out.append(None)
        else:
            out.append(patterns_from_tuples(cut, seq, tuples[offsets[g]:offsets[g + 1]], length_weight))
    return out


def top_patterns_batch(rel_cut_positions, sequences, length_weight, n_patterns, device=None, stats=None):
    """The ``n_patterns`` best-scoring patterns per cut site, in the order
    ``sorted(patterns, key=lambda x: -x["pattern_score"])[:n]`` gives (stable), or None.

    Same dicts as :func:`generate_mmej_patterns_batch` would give, but only the winners are built:
    the scores of ALL kept candidates are evaluated vectorised -- ``100 * lf[del_len] * (len + GC)``
    in float64 with ``lf`` tabulated by the reference's own expression, i.e. the same two IEEE
    multiplications in the same order, hence bit-identical -- and the stable top-n is taken on them.
    (A 150-bp window keeps ~240 patterns; guido keeps 5.)
    """
    sequences = list(sequences)
    cuts = _split_points(rel_cut_positions, sequences)
    if not sequences:
        return []
    tuples, offsets, n_cand = _native.mmej_batch(sequences, cuts, device=device, stats=stats)
    winners = _top_indices(tuples, offsets, sequences, cuts, length_weight, n_patterns)
    out = []
    for g, (seq, cut) in enumerate(zip(sequences, cuts)):
        if n_cand[g] == 0:
            out.append(None)
            continue
        left_seq, right_seq = seq[:cut], seq[cut:]
        ls, le, rs = _unpack(tuples[winners[g]])
        out.append([_pattern_dict(left_seq, right_seq, a, b, c, length_weight) for a, b, c in zip(ls, le, rs)])
    return out


def _top_indices(tuples, offsets, sequences, cuts, length_weight, n_patterns):
    """Per window, the indices (into ``tuples``) of its ``n_patterns`` best candidates in the order
    ``sorted(key=-score)`` gives (stable: ties keep candidate order).  All windows at once: the scores
    are one vectorised expression over every kept candidate -- the reference's, term by term, in
    float64 -- and the winners are taken by n_patterns rounds of "first maximum of every segment"."""
    n_win = len(sequences)
    offsets = np.asarray(offsets, dtype=np.int64)
    counts = np.diff(offsets)
    total = int(offsets[-1])
    if total == 0 or n_patterns <= 0:
        return [np.zeros(0, np.int64) for _ in range(n_win)]
    ls, le, rs = _unpack(tuples[:total])
    g = np.repeat(np.arange(n_win), counts)
    left_len = np.asarray(cuts, dtype=np.int64)
    # G/C prefix sums over the concatenated left flanks
    cat = np.frombuffer("".join(seq[:cut] for seq, cut in zip(sequences, cuts)).encode(), dtype=np.uint8)
    cum = np.concatenate(([0], np.cumsum((cat == 71) | (cat == 67))))
    loff = np.concatenate(([0], np.cumsum(left_len)))[:-1]
    gc = cum[loff[g] + le] - cum[loff[g] + ls]
    del_len = (left_len[g] - ls) + rs
    lf = np.array([round(1 / math.exp(d / length_weight), 3) for d in range(int(del_len.max()) + 1)], dtype=np.float64)
    work = (100 * lf[del_len]) * (((le - ls) - gc) + (gc * 2))
    nonempty = np.nonzero(counts > 0)[0]
    seg_starts = offsets[:-1][nonempty]
    seg_counts = counts[nonempty]
    position = np.arange(total)
    picks = np.full((len(nonempty), n_patterns), -1, dtype=np.int64)
    for k in range(n_patterns):
        seg_max = np.maximum.reduceat(work, seg_starts)
        first = np.minimum.reduceat(np.where(work == np.repeat(seg_max, seg_counts), position, total), seg_starts)
        live = seg_max != -np.inf            # windows with fewer than k + 1 candidates are exhausted
        picks[live, k] = first[live]
        work[first[live]] = -np.inf
        if not live.any():
            break
    out = [np.zeros(0, np.int64) for _ in range(n_win)]
    for row, w in zip(picks, nonempty.tolist()):
        out[w] = row[row >= 0]
    return out


def generate_mmej_patterns(rel_cut_position, sequence, length_weight):
    """All MMEJ patterns around ``rel_cut_position`` of ``sequence`` (list of dicts or None)."""
    return generate_mmej_patterns_batch([rel_cut_position], [sequence], length_weight)[0]
