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


def test_mmej_general_kernel_equals_fast(native, oracle, monkeypatch):
    """the bit-parallel kernel (default) and the general byte-compare kernel give the same tuples"""
    rng = np.random.default_rng(5)
    windows = [util.random_dna(rng, int(n)) for n in rng.integers(20, 255, size=150)] + ["ACGT" * 30, "AAC" * 60]
    cuts = [len(w) // 2 if i % 3 else min(len(w), 7 * i % 120) for i, w in enumerate(windows)]
    fast = native.mmej_batch(windows, cuts)
    monkeypatch.setenv("GUIDO_MMEJ_GENERAL", "1")
    general = native.mmej_batch(windows, cuts)
    for x, y in zip(fast, general):
        assert np.array_equal(x, y)
    _check(native, oracle, windows[:40], cuts[:40])


def test_mmej_sites_on_the_resident_image(native, oracle):
    """windows addressed by position in the uploaded image == the same windows sent as letters,
    including windows over N runs / IUPAC letters (host fetch + general kernel) and clipped ones"""
    contigs = util.synthetic_contigs(77, [30000, 8000], n_runs=6)
    img = native.Image.from_seqs(contigs).upload()
    rng = np.random.default_rng(6)
    (_, l0, o0), (_, l1, o1) = img.contigs[0][:3], img.contigs[1][:3]
    starts = np.concatenate([o0 + rng.integers(0, l0 - 160, size=300), o1 + rng.integers(0, l1 - 160, size=60),
                             [o0, o0 + l0 - 150, o1 + l1 - 90]]).astype(np.int64)
    lens = np.concatenate([np.full(360, 150), [150, 150, 90]]).astype(np.int32)
    cuts = np.concatenate([np.full(300, 75), rng.integers(0, 150, size=60), [75, 75, 75]]).astype(np.int32)
    # runs of the synthetic contigs: make sure some windows cover them
    S0 = contigs[0][1].upper()
    bad = [i for i, ch in enumerate(S0) if ch not in "ACGT"][:3]
    for b in bad:
        starts = np.append(starts, o0 + max(b - 70, 0)); lens = np.append(lens, 150); cuts = np.append(cuts, 75)
    t1, off1, nc1 = img.mmej_sites(starts, lens, cuts)
    windows = []
    for s0, ln in zip(starts, lens):
        windows.append(bytes(img.fetch_windows([s0], int(ln))[0]).decode())
    t2, off2, nc2 = native.mmej_batch(windows, cuts.tolist())
    assert np.array_equal(off1, off2) and np.array_equal(nc1, nc2) and np.array_equal(t1, t2)
    assert any(set(w) - set("ACGT") for w in windows) or not bad
    _check(native, oracle, windows[:25], cuts[:25].tolist())


def test_mmej_cut_positions_follow_slice_semantics(native):
    """a negative or out-of-range cut splits the sequence where Python's sequence[:cut] / sequence[cut:]
    do (reference mmej.py:13-14), in both batch entry points"""
    from guido_b200 import mmej
    rng = np.random.default_rng(8)
    seq = util.random_dna(rng, 140)
    for cut, same_as in ((-30, 110), (-500, 0), (500, 140)):
        assert mmej.generate_mmej_patterns(cut, seq, 20) == mmej.generate_mmej_patterns(same_as, seq, 20)
        assert mmej.top_patterns_batch([cut], [seq], 20, 5) == mmej.top_patterns_batch([same_as], [seq], 20, 5)
    assert mmej.generate_mmej_patterns(-30, seq, 20) is not None


def test_locus_layer_methods(native):
    import guido_b200 as guido
    rng = np.random.default_rng(9)
    loc = guido.Locus(sequence=util.random_dna(rng, 400), name="c", start=101)
    loc.find_guides()
    data = np.arange(400)
    loc.add_layer("pos", data)
    g = loc.guides[0]
    assert np.array_equal(g.layer("pos"), data[g.guide_start - 101:g.guide_start - 101 + 23])
    with pytest.raises(ValueError):
        loc.add_layer("short", np.arange(10))
    with pytest.raises(ValueError):
        loc.guides_to_csv("")
    with pytest.raises(NotImplementedError):
        loc.rank_guides()
    assert callable(loc.add_azimuth_score) and callable(loc.guides_detailed_table)


def test_find_guides_with_a_pam_longer_than_the_matcher(native):
    """a PAM of more than 8 letters gives no guides (the reference keeps 23-letter guides only) instead of
    an error from the C layer; an unknown letter is still a KeyError"""
    import guido_b200 as guido
    rng = np.random.default_rng(10)
    loc = guido.Locus(sequence=util.random_dna(rng, 300), name="c", start=1)
    assert loc.find_guides(pam="NGGNGGNGG") == [] or len(loc.guides) == 0
    with pytest.raises(KeyError):
        loc.find_guides(pam="NGGNGGNGGZ")
