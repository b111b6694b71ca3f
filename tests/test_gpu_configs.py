"""The BASELINE.json configurations as parity tests on a real B200.

Small sizes are compared with the CPU oracle record for record; the full sizes (where a brute
force oracle would take minutes to hours) are checked through size-independent properties:
two independent GPU algorithms agreeing byte for byte, every hit re-verified on the host from the
decoded genome, self-hits present, sortedness, counts against a numpy recount of a slice."""
import numpy as np
import pytest

import guido_b200 as guido
from guido_b200 import _native
from guido_b200.off_targets import decode_hits
from tests import util

pytestmark = pytest.mark.gpu


def _guides_from_sites(img, n, seed):
    import bench
    return bench.sample_guides(img, n, seed)


def _verify_hits(img, hits, guides_u8, total_mm=4, core_len=10, core_mm=2):
    """host re-check of EVERY hit: PAM is NGG, mismatch budget respected, payload = genome"""
    w = decode_hits(hits, 23)
    assert ((w[:, 21] == ord("G")) & (w[:, 22] == ord("G"))).all()
    diff = w[:, :20] != guides_u8[hits["guide"]]
    assert (diff.sum(axis=1) <= total_mm).all()
    assert (diff[:, 20 - core_len:].sum(axis=1) <= core_mm).all()
    # the payload really is the genome at gpos (forward strand), reverse-complemented for '-'
    fwd = img.fetch_windows(hits["gpos"].astype(np.int64), 23)
    minus = (hits["hi"] >> 31).astype(bool)
    comp = np.zeros(256, np.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    rc = comp[fwd[:, ::-1]]
    assert np.array_equal(np.where(minus[:, None], rc, fwd), w)


def test_config1_scaled_matches_oracle(tmp_path, oracle):
    """find_guides + off-targets (<=3 mm) on a locus of a synthetic genome, vs the oracle"""
    contigs = util.synthetic_contigs(1001, [400_000, 350_000, 250_000], n_runs=5, run_len=(50, 2000), lower_frac=0.1)
    fa = tmp_path / "g.fa"
    with open(fa, "w") as fh:
        for name, s in contigs:
            fh.write(f">{name}\n")
            fh.writelines(s[i:i + 60] + "\n" for i in range(0, len(s), 60))
    g = guido.Genome("g", genome_file_abspath=str(fa))
    g.build()
    loc = guido.locus_from_coordinates(g, "chr2", 100_001, 104_000)
    guides = loc.find_guides()
    want = oracle.find_guides_records(loc.sequence.seq, 100_001, 104_000, chromosome="chr2")
    assert [(x.guide_seq, x.guide_start, x.guide_strand, x.absolute_cut_pos) for x in guides] == \
        [(w["guide_seq"], w["guide_start"], w["guide_strand"], w["absolute_cut_pos"]) for w in want]
    found = loc.find_off_targets(total_mismatches=3)
    exp = oracle.off_target_records(want, contigs, total_mismatches=3)
    assert set(found) == set(exp) == set(range(len(guides)))  # the self-hit gives every guide a key
    key = lambda r: (r["chromosome"], r["start"], r["strand"])
    for ix in exp:
        got = sorted(({k: v for k, v in r.items() if k != "cfd_score"} for r in found[ix]), key=key)
        assert got == exp[ix]


def test_config1_full_size_properties():
    """10 Mb genome, 100 kb locus (~12.5 k guides), <=3 mm: index kernel == brute-force kernel,
    every hit re-verified, every guide finds itself"""
    img = _native.Image.synth([2_500_000] * 4, seed=1001, n_frac=0.005).upload()
    seq = img.fetch(1, 1_000_000, 100_000)
    loc = guido.Locus(sequence=seq, name="chr2", start=1_000_001)
    guides = loc.find_guides()
    assert 10_500 < len(guides) < 14_000
    protos = np.frombuffer("".join(x.guide_seq[:20] for x in guides).encode(), np.uint8).reshape(-1, 20)
    a, fa = img.offtarget_search(protos, total_mismatches=3, algo=_native.OT_INDEX)
    b, fb = img.offtarget_search(protos, total_mismatches=3, algo=_native.OT_SCAN)
    assert np.array_equal(a, b) and np.array_equal(fa, fb) and fa.all()
    _verify_hits(img, a, protos, total_mm=3)
    w = decode_hits(a, 20)
    zero = (w == protos[a["guide"]]).all(axis=1)
    assert set(a["guide"][zero].tolist()) == set(range(len(guides)))


def test_config2_whole_genome_scan_properties():
    """273 Mb NGG scan: ascending, PAM present at every sampled position, exact agreement with a
    numpy recount of one whole contig, the scan modes nest"""
    import bench
    img = _native.Image.synth(bench.AGAM_LENS, seed=1002, n_frac=0.005).upload()
    fw, rv = img.pam_scan_genome("NGG", _native.SCAN_GUIDE)
    assert (np.diff(fw.astype(np.int64)) > 0).all() and (np.diff(rv.astype(np.int64)) > 0).all()
    assert abs(len(fw) / 273.2e6 - 1 / 16) < 0.002 and abs(len(rv) / 273.2e6 - 1 / 16) < 0.002
    rng = np.random.default_rng(0)
    for arr, sl, pat in ((fw, slice(1, 3), b"GG"), (rv, slice(0, 2), b"CC")):
        pick = arr[rng.integers(0, len(arr), 20000)]
        w = img.fetch_windows(pick.astype(np.int64), 3)
        assert (w[:, sl] == np.frombuffer(pat, np.uint8)).all()
    name, clen, goff = img.contigs[4]  # 24.4 Mb
    s = np.frombuffer(img.fetch(4, 0, clen).encode(), np.uint8)
    isg, isc = s == ord("G"), s == ord("C")
    bad = np.convolve((s == ord("N")).astype(np.int32), np.ones(23, np.int32))  # bad[k] = #N in s[k-22 .. k]
    p = np.arange(20, clen - 2)
    ok_f = isg[p + 1] & isg[p + 2] & (bad[p + 2] == 0)
    q = np.arange(0, clen - 22)
    ok_r = isc[q] & isc[q + 1] & (bad[q + 22] == 0)
    in_f = fw[(fw >= goff) & (fw < goff + clen)] - goff
    in_r = rv[(rv >= goff) & (rv < goff + clen)] - goff
    assert np.array_equal(in_f, p[ok_f]) and np.array_equal(in_r, q[ok_r])
    f0, r0 = img.pam_scan_genome("NGG", _native.SCAN_REGEX)
    f2, r2 = img.pam_scan_genome("NGG", _native.SCAN_STRICT)
    assert len(f0) >= len(fw) == len(f2) and len(r0) >= len(rv) == len(r2)  # synthetic: only N runs and pads


def test_config3_10k_guides_properties():
    """10 k guides, <=4 mm vs 273 Mb: every hit re-verified on the host, self-hits complete,
    the brute-force kernel agrees on one whole contig"""
    import bench
    img = _native.Image.synth(bench.AGAM_LENS, seed=1002, n_frac=0.005).upload()
    protos = _guides_from_sites(img, 10_000, 1003)
    hits, flags = img.offtarget_search(protos)
    assert flags.all() and 80_000 < len(hits) < 140_000
    assert (np.diff(hits["guide"].astype(np.int64)) >= 0).all()
    _verify_hits(img, hits, protos)
    w = decode_hits(hits, 20)
    zero = (w == protos[hits["guide"]]).all(axis=1)
    assert len(set(hits["guide"][zero].tolist())) == 10_000
    name, clen, goff = img.contigs[4]
    sub = _native.Image.from_seqs([(name, img.fetch(4, 0, clen))]).upload()
    bh, _ = sub.offtarget_search(protos[:1500], algo=_native.OT_SCAN, want_raw_flags=False)
    ih = hits[(hits["guide"] < 1500) & (hits["gpos"] >= goff) & (hits["gpos"] < goff + clen)]
    assert len(bh) == len(ih) and len(bh) > 0
    assert np.array_equal(bh["gpos"] - sub.contigs[0][2], ih["gpos"] - goff)
    assert np.array_equal(bh["hi"], ih["hi"]) and np.array_equal(bh["lo"], ih["lo"])


def test_config4_full_size_properties():
    """100 k guides, <=4 mm vs the 3.1 Gb genome (the bench workload) at full size: the table join
    (bit-sliced compare) and the streaming index kernel return the same hit list byte for byte,
    every guide finds the site it was taken from, hits come back in canonical order, a sample of
    them is re-verified on the host, and two key-range parts of the site table add up to the whole"""
    import bench
    lens, seed, n_frac, _ = bench.genome_lens("human")
    img = _native.Image.synth(lens, seed=seed, n_frac=n_frac).upload()
    protos = bench.sample_guides(img, 100_000, seed + 1)
    hits, _ = img.offtarget_search(protos, algo=_native.OT_TABLE, want_raw_flags=False, pinned=True)
    hits = hits.copy()
    assert 9_000_000 < len(hits) < 13_000_000
    key = (hits["guide"].astype(np.uint64) << np.uint64(33)) | (hits["gpos"].astype(np.uint64) << np.uint64(1)) \
        | (hits["hi"] >> 31).astype(np.uint64)
    assert (np.diff(key.astype(np.int64)) > 0).all()            # canonical order, no duplicates
    enc = _native.encode_guides([bytes(p).decode() for p in protos]).reshape(-1, 2)   # (hi, lo) planes
    m20 = np.uint32((1 << 20) - 1)
    zero = ((hits["hi"] & m20) == enc[hits["guide"], 0]) & ((hits["lo"] & m20) == enc[hits["guide"], 1])
    assert len(np.unique(hits["guide"][zero])) == 100_000        # every guide finds its own site
    rng = np.random.default_rng(4)
    _verify_hits(img, hits[np.sort(rng.choice(len(hits), 200_000, replace=False))], protos)
    streamed, _ = img.offtarget_search(protos, algo=_native.OT_INDEX, want_raw_flags=False, pinned=True)
    assert np.array_equal(streamed, hits)
    del streamed
    n_parts, total = 2, 0
    for p in range(n_parts):
        img.site_table_partition(p, n_parts)
        part, _ = img.offtarget_search(protos, algo=_native.OT_TABLE, want_raw_flags=False, pinned=True)
        part_key = (part["guide"].astype(np.uint64) << np.uint64(33)) | (part["gpos"].astype(np.uint64) << np.uint64(1)) \
            | (part["hi"] >> 31).astype(np.uint64)
        assert np.isin(part_key, key, assume_unique=True).all()
        total += len(part)
    assert total == len(hits)


def test_config5_tttv_scan_and_mmej(oracle):
    """Cas12a TTTV scan + MMEJ for the guides of an exon set (scaled: 40 genes)"""
    contigs = util.synthetic_contigs(1006, [600_000], n_runs=3, run_len=(50, 500))
    seq = contigs[0][1]
    rng = np.random.default_rng(1006)
    windows, cuts = [], []
    n_tttv = 0
    for gene in range(40):
        start = int(rng.integers(1000, 590_000))
        for exon in range(int(rng.integers(3, 9))):
            a = start + exon * 700
            b = a + int(rng.integers(100, 500))
            exon_seq = seq[a:b]
            loc = guido.Locus(sequence=exon_seq, name="chr1", start=a + 1)
            got = loc._find_guides_in_interval(exon_seq, a + 1, "TTTV")
            want = oracle.guide_records(exon_seq, a + 1, "TTTV", chromosome="chr1")
            assert [(x.guide_seq, x.guide_start, x.guide_strand) for x in got] == \
                [(w["guide_seq"], w["guide_start"], w["guide_strand"]) for w in want]
            assert loc.find_guides(pam="TTTV") == []  # upstream: only 3-nt PAMs give 23-mers (locus.py:316)
            n_tttv += len(got)
            for x in loc.find_guides(pam="NGG"):
                if "N" not in x.locus_seq:
                    windows.append(x.locus_seq)
                    cuts.append(x.relative_cut_pos)
    assert n_tttv > 300 and len(windows) > 3000
    from guido_b200.mmej import generate_mmej_patterns_batch
    got = generate_mmej_patterns_batch(cuts[:1500], windows[:1500], 20)
    for w, c, pats in zip(windows[:1500:7], cuts[:1500:7], got[::7]):
        assert pats == oracle.mmej_patterns(c, w, 20)
