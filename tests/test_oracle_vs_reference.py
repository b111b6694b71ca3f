"""Differential tests: the CPU oracle against the LIVE reference (imported from /root/reference via
oracle/ref_loader.py).  Build container only -- skipped where the reference is not mounted."""
import numpy as np
import pytest

from oracle import fake_bowtie, ref_loader
from tests import util

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")]

ATTRS = ["guide_seq", "guide_long_seq", "guide_pam", "guide_chrom", "guide_start", "guide_end",
         "guide_strand", "relative_cut_pos", "absolute_cut_pos", "locus_seq"]


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load_reference()


@pytest.mark.parametrize("seed", range(8))
def test_find_guides_differential(ref, oracle, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(30, 3000))
    pam = ["NGG", "NAG", "TTTV", "NNGRRT", "YG", "NGN", "NNN", "GGG"][seed]
    flank = int(rng.integers(0, 100))
    start = int(rng.integers(1, 10**6))
    seq = util.synthetic_contigs(seed + 500, [n], n_runs=2, run_len=(1, 30), lower_frac=0.3, iupac=seed % 3)[0][1]
    loc = ref.Locus(sequence=seq, name="chrQ", start=start)
    want = loc.find_guides(pam=pam, min_flanking_length=flank)
    mine = oracle.find_guides_records(seq, start, None, pam=pam, min_flanking_length=flank, chromosome="chrQ")
    assert len(mine) == len(want)
    for a, b in zip(mine, want):
        assert a["id"] == b.id
        for k in ATTRS:
            assert a[k] == getattr(b, k), k
    raw = loc._find_guides_in_interval(seq, start, pam)
    mine_raw = oracle.guide_records(seq, start, pam, chromosome="chrQ")
    assert [(g.guide_seq, g.guide_start, g.guide_strand) for g in raw] == \
        [(g["guide_seq"], g["guide_start"], g["guide_strand"]) for g in mine_raw]


@pytest.mark.parametrize("seed", range(6))
def test_mmej_differential(ref, oracle, seed):
    rng = np.random.default_rng(seed + 900)
    for _ in range(6):
        n = int(rng.integers(4, 160))
        cut = int(rng.integers(0, n + 1))
        alphabet = "ACGT"[: int(rng.integers(1, 5))]
        seq = "".join(rng.choice(list(alphabet), n))
        lw = float(rng.choice([20, 10, 33.3]))
        assert oracle.mmej_patterns(cut, seq, lw) == ref.mmej.generate_mmej_patterns(cut, seq, lw), (seq, cut)


@pytest.mark.parametrize("pam,kw", [("NGG", {}), ("NGG", dict(core_mismatches=1, total_mismatches=2)),
                                    ("NRG", dict(core_mismatches=1)), ("TGG", dict(core_mismatches=3))])
def test_offtarget_postprocessing_differential(ref, oracle, pam, kw):
    """reference run_bowtie (fed by the fake bowtie) == oracle.off_target_records on the same raw hits"""
    contigs = util.synthetic_contigs(77, [20000, 15000], n_runs=2)
    s0 = contigs[0][1]
    contigs[1] = (contigs[1][0], contigs[1][1][:3000] + s0[2000:2400].upper() + contigs[1][1][3400:])
    index = "/virtual/idx"
    loc = ref.Locus(sequence=s0[1990:2410], name="chr1", start=1991, genome=fake_bowtie.FakeGenome(index))
    loc.find_guides(pam=pam)
    guides = oracle.find_guides_records(s0[1990:2410], 1991, None, pam=pam, chromosome="chr1")
    if "N" not in pam.replace("R", "N"):
        # upstream quirk: a PAM without N makes the self-hit a perfect alignment, whose empty
        # descriptor field crashes _parse_mismatches (off_targets.py:73); the oracle mirrors it
        with fake_bowtie.patched(ref, {index: contigs}):
            with pytest.raises(ValueError, match="not enough values"):
                ref.off_targets.run_bowtie(loc.guides, pam=pam, genome_index_path=index, **kw)
        with pytest.raises(ValueError, match="not enough values"):
            oracle.off_target_records(guides, contigs, pam=pam, **kw)
        return
    with fake_bowtie.patched(ref, {index: contigs}):
        want = ref.off_targets.run_bowtie(loc.guides, pam=pam, genome_index_path=index, **kw)
    mine = oracle.off_target_records(guides, contigs, pam=pam, **kw)
    assert set(mine) == set(want)
    key = lambda r: (r["chromosome"], r["start"], r["strand"])
    for ix in mine:
        assert sorted(want[ix], key=key) == mine[ix]
    assert sum(len(v) for v in mine.values()) > 0


@pytest.mark.parametrize("seed", range(6))
def test_package_host_logic_differential(ref, monkeypatch, seed):
    """guido_b200's host side (Guide geometry, interval handling, sorting / naming, MMEJ pattern dicts, scores,
    top-n and the out-of-frame layer) with the device entry points replaced by the oracle-backed fakes ==
    the live reference on the same random inputs"""
    import guido_b200 as guido
    from guido_b200 import _native
    from tests import fakes
    monkeypatch.setattr(_native, "pam_scan_seq", fakes.fake_pam_scan_seq)
    monkeypatch.setattr(_native, "mmej_batch", fakes.fake_mmej_batch)
    rng = np.random.default_rng(seed + 4000)
    n = int(rng.integers(60, 1500))
    pam = ["NGG", "NAG", "NGN", "NNN", "YGG", "NGG"][seed]
    flank = int(rng.integers(0, 60))
    start = int(rng.integers(1, 10**5))
    seq = util.synthetic_contigs(seed + 700, [n], n_runs=1, run_len=(1, 10), lower_frac=0.2)[0][1]
    want_loc = ref.Locus(sequence=seq, name="chrQ", start=start)
    mine_loc = guido.Locus(sequence=seq, name="chrQ", start=start)
    want = want_loc.find_guides(pam=pam, min_flanking_length=flank)
    mine = mine_loc.find_guides(pam=pam, min_flanking_length=flank)
    assert len(mine) == len(want)
    for a, b in zip(mine, want):
        assert a.id == b.id
        for k in ATTRS:
            assert getattr(a, k) == getattr(b, k), k
    if not want:
        return
    n_top, lw = int(rng.integers(1, 8)), float(rng.choice([20, 10]))
    try:
        want_loc.simulate_end_joining(n_patterns=n_top, length_weight=lw)
        failed = None
    except TypeError as exc:      # upstream: a guide without any MMEJ pattern
        failed = exc
    if failed is not None:
        with pytest.raises(TypeError):
            mine_loc.simulate_end_joining(n_patterns=n_top, length_weight=lw)
        return
    mine_loc.simulate_end_joining(n_patterns=n_top, length_weight=lw)
    for a, b in zip(mine_loc.guides, want_loc.guides):
        assert a.mmej_patterns == b.mmej_patterns
        assert a.layers == b.layers
