"""GPU parity of the MMEJ warp kernel against the CPU oracle (C restatement of guido/mmej.py)."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


def _unpack(t):
    t = np.asarray(t, dtype=np.uint32)
    return np.stack([t & 255, (t >> 8) & 255, (t >> 16) & 255, t >> 24], axis=1).astype(np.int64)


def _check(native, oracle, windows, cuts):
    tuples, offs, ncand = native.mmej_batch(windows, cuts)
    for g, (w, c) in enumerate(zip(windows, cuts)):
        kept, nc = oracle.mmej_tuples(w, c)
        got = _unpack(tuples[offs[g]:offs[g + 1]])
        if kept is None:
            assert ncand[g] == 0 and len(got) == 0, (g, w, c)
        else:
            assert ncand[g] == nc, (g, w, c)
            assert np.array_equal(got, kept), (g, w, c)


def test_mmej_random_windows(native, oracle):
    rng = np.random.default_rng(0)
    windows = [util.random_dna(rng, 150) for _ in range(200)]
    _check(native, oracle, windows, [75] * 200)


def test_mmej_ragged_and_clipped(native, oracle):
    rng = np.random.default_rng(1)
    windows, cuts = [], []
    for n in [0, 1, 2, 3, 4, 5, 10, 40, 76, 100, 149, 150, 151, 200, 300, 510]:
        for cut in sorted({0, 1, 2, 3, n // 2, max(n - 2, 0), n}):
            if cut <= 255 and n - cut <= 255:
                windows.append(util.random_dna(rng, n))
                cuts.append(cut)
    _check(native, oracle, windows, cuts)


def test_mmej_low_complexity(native, oracle):
    windows = ["A" * 150, "AT" * 75, "ACG" * 50, "AAAAACCCCC" * 15, "ACGT" * 37 + "AC",
               "A" * 75 + "C" * 75, "GATTACA" * 21 + "GAT", "N" * 150, "ACGTN" * 30]
    _check(native, oracle, windows, [75] * len(windows))
    _check(native, oracle, windows, [50] * len(windows))


def test_mmej_very_repetitive_needs_retry(native, oracle):
    # > 8192 candidates in one window: exercises the scratch-overflow retry pass
    windows = ["AC" * 255, "A" * 400, util.random_dna(np.random.default_rng(2), 150)]
    cuts = [255, 200, 75]
    _check(native, oracle, windows, cuts)


def test_mmej_limits(native):
    with pytest.raises(ValueError):
        native.mmej_batch(["A" * 600], [300])
    t, offs, nc = native.mmej_batch([], [])
    assert len(t) == 0 and offs.tolist() == [0]
