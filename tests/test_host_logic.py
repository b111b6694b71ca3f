"""Host-side logic of the package (record building, filters, I/O conventions) on a box WITHOUT a GPU.

The three device entry points are replaced by oracle-backed fakes (tests/fakes.py); everything
above the C-ABI is the shipped code and is compared with the golden fixtures generated from the
reference.  The same comparisons run against the real kernels in tests/test_gpu_api.py.
"""
import math
import os
import pickle

import numpy as np
import pandas as pd
import pytest

import guido_b200 as guido
from guido_b200 import _native, off_targets
from tests import fakes, util

GUIDE_ATTRS = ["id", "guide_seq", "guide_long_seq", "guide_pam", "guide_chrom", "guide_start", "guide_end",
               "guide_strand", "relative_cut_pos", "absolute_cut_pos", "locus_seq"]


@pytest.fixture
def cpu_fakes(monkeypatch):
    monkeypatch.setattr(_native, "pam_scan_seq", fakes.fake_pam_scan_seq)
    monkeypatch.setattr(_native, "mmej_batch", fakes.fake_mmej_batch)


def same_guides(guides, golden, attrs=GUIDE_ATTRS):
    assert len(guides) == len(golden)
    for g, b in zip(guides, golden):
        for k in attrs:
            assert getattr(g, k) == b[k], (k, getattr(g, k), b[k])


def same_layers(mine, golden):
    assert set(mine) == set(golden)
    for k, v in golden.items():
        if isinstance(v, float) and math.isnan(v):
            assert math.isnan(mine[k]), k
        else:
            assert mine[k] == v, (k, mine[k], v)


def test_kat_loci(cpu_fakes):
    kat = fakes.load_golden("kat.json")
    for key in ("locus101", "locus126", "locus490", "agap005958"):
        c = kat[key]
        loc = guido.Locus(sequence=c["seq"], name=c["name"], start=c["start"], end=c.get("end"))
        same_guides(loc.find_guides(), c["guides"])
    assert repr(loc.guides[0]).startswith("gRNA-1(")
    assert loc.guide("gRNA-2") is loc.guides[1] and loc.guide(0) is loc.guides[0]
    with pytest.raises(ValueError):
        loc.guide("nope")


def test_kat_mmej(cpu_fakes):
    kat = fakes.load_golden("kat.json")["agap005958"]
    loc = guido.Locus(sequence=kat["seq"], name=kat["name"], start=kat["start"], end=kat["end"])
    loc.find_guides()
    loc.simulate_end_joining()
    assert loc.guide(5).mmej_patterns == kat["stored_mmej_guide5"]
    for g, m in zip(loc.guides, kat["mmej"]):
        assert g.mmej_patterns == m["patterns"] and g.mmej_str == m["mmej_str"]
        same_layers(g.layers, m["layers"])
    # per-guide entry point gives the same answer as the batched one
    g = loc.guides[7]
    before = (g.mmej_patterns, dict(g.layers))
    g.simulate_end_joining()
    assert (g.mmej_patterns, dict(g.layers)) == before


def test_synthetic_find_guides(cpu_fakes):
    for c in fakes.load_golden("guides_synth.json.gz"):
        ann = pd.DataFrame(c["annotation"]) if c["annotation"] else None
        loc = guido.Locus(sequence=c["seq"], name="chrS", start=c["start"], annotation=ann)
        guides = loc.find_guides(pam=c["pam"], min_flanking_length=c["min_flanking_length"],
                                 selected_features=c["selected_features"] or "all")
        assert [list(map(int, i)) for i in loc.intervals] == c["intervals"]
        same_guides(guides, c["guides"])
        raw = loc._find_guides_in_interval(c["seq"][:c["interval_prefix"]], c["start"], c["pam"])
        same_guides(raw, c["interval_guides"], [a for a in GUIDE_ATTRS if a != "id"])


def test_guide_constructor_matches_fast_path(cpu_fakes):
    seq = util.synthetic_contigs(3, [500], n_runs=1)[0][1].upper()
    loc = guido.Locus(sequence=seq, name="c", start=100)
    for g in loc._find_guides_in_interval(seq, 100, "NGG"):
        p = (g.absolute_cut_pos - 100 + 3) if g.guide_strand == "+" else (g.absolute_cut_pos - 100 - 6)
        h = guido.Guide(seq, p, 3, strand=g.guide_strand, chromosome="c", start=100)
        assert h.__dict__ == g.__dict__
    with pytest.raises(TypeError):
        g.no_such_attribute
    with pytest.raises(ValueError):
        g.layer("missing")


def test_synthetic_mmej(cpu_fakes):
    from guido_b200.mmej import generate_mmej_patterns
    for c in fakes.load_golden("mmej_synth.json.gz"):
        got = generate_mmej_patterns(c["cut"], c["seq"], c["length_weight"])
        assert got == c["patterns"]
        if got:
            assert isinstance(got[0]["left_seq_position"], np.int64)  # as the reference's numpy rows


def test_mmej_none_raises_like_reference(cpu_fakes):
    loc = guido.Locus(sequence="ACGT" * 5 + "TGG" + "ACGT" * 5, name="c")
    g = guido.Guide("A" * 10 + "C" * 10, 10, 3)
    g.locus_seq, g.relative_cut_pos = "AAAACCCC", 4
    with pytest.raises(TypeError):
        g.simulate_end_joining()
    with pytest.raises(ValueError):
        guido.Locus(sequence="ACGTACGT", name="c").simulate_end_joining()


def test_offtarget_records(cpu_fakes, monkeypatch):
    for c in fakes.load_golden("offtargets_synth.json.gz"):
        contigs = [tuple(x) for x in c["contigs"]]
        image = fakes.FakeImage(contigs)
        monkeypatch.setattr(off_targets, "get_image", lambda path, device=None, **kw: image)
        for spec, golden, layers in ((c["locus"], c["result"], c["layers"]),
                                     (c["external"], c["external"]["result"], c["external"]["layers"])):
            genome = type("G", (), {"bowtie_index": "/virtual/index"})()
            loc = guido.Locus(sequence=spec["seq"], name=spec["name"], start=spec["start"], genome=genome)
            loc.find_guides(pam=c["pam"])
            found = loc.find_off_targets(**c["kwargs"])
            golden = fakes.int_keys(golden)
            assert set(found) == set(golden)
            for ix, recs in golden.items():
                want = sorted((dict(r, mismatches=fakes.int_keys(r["mismatches"])) for r in recs),
                              key=lambda r: (r["chromosome"], r["start"], r["strand"]))
                got = sorted(found[ix], key=lambda r: (r["chromosome"], r["start"], r["strand"]))
                assert got == want
            for g, lay in zip(loc.guides, layers):
                if "ot_cfd_score_mean" in lay and not math.isnan(lay["ot_cfd_score_mean"]):
                    # mean/sum depend on bowtie's (unspecified) emission order only in the last ulp
                    assert g.layers["ot_sum_score"] == lay["ot_sum_score"]
                    assert g.layers["ot_cfd_score_max"] == lay["ot_cfd_score_max"]
                    assert math.isclose(g.layers["ot_cfd_score_sum"], lay["ot_cfd_score_sum"], rel_tol=1e-12)
                    assert math.isclose(g.layers["ot_cfd_score_mean"], lay["ot_cfd_score_mean"], rel_tol=1e-12)
                else:
                    same_layers(g.layers, lay)
            if spec is c["locus"]:
                assert [g.off_target_str for g in loc.guides] == c["off_target_str"]
                assert [g.off_targets_string for g in loc.guides] == c["off_targets_string"]


def test_offtarget_argument_errors(cpu_fakes, monkeypatch):
    loc = guido.Locus(sequence="ACGT" * 30, name="c")
    with pytest.raises(ValueError, match="No genome"):
        loc.find_off_targets()
    loc.genome = type("G", (), {"bowtie_index": None})()
    with pytest.raises(ValueError, match="not built"):
        loc.find_off_targets()


def test_rev_comp_semantics():
    assert guido.rev_comp("ACGTNRX") == "NNNACGT"
    assert guido.rev_comp("acgt") == "NNNN"
    assert guido.rev_comp("ACGU-", rna=True) == "-ACGT"
    assert guido.rev_comp("N", rna=True) == "N"  # not in the RNA table -> 'N' by the fallback
    from guido_b200.helpers import rev_comp_fast
    for s in ("ACGTNRYKMacgtn-*", "", "G"):
        assert rev_comp_fast(s) == guido.rev_comp(s)


# ------------------------------------------------------------------ Genome.build / factories

def _write_genome(tmp_path, gtf=True):
    contigs = util.synthetic_contigs(31, [4000, 2500], n_runs=2, run_len=(5, 60))
    fa = tmp_path / "toy.fa"
    with open(fa, "w") as fh:
        for name, s in contigs:
            fh.write(f">{name} description text\n")
            for i in range(0, len(s), 60):
                fh.write(s[i:i + 60] + "\n")
    ann = None
    if gtf:
        ann = tmp_path / "toy.gff3"
        rows = [("chr1", "src", "gene", 1001, 2000, ".", "+", ".", "ID=GENE1;Name=g1"),
                ("chr1", "src", "exon", 1001, 1200, ".", "+", ".", "ID=e1;Parent=GENE1;Name=E1;gene_id=GENE1"),
                ("chr1", "src", "exon", 1500, 2000, ".", "+", ".", "ID=e2;Parent=GENE1;Name=E2;gene_id=GENE1"),
                ("chr2", "src", "gene", 100, 900, ".", "-", ".", "ID=GENE2;Name=g2"),
                ("chr1", "src", "gene", 10, 400, ".", "-", ".", "ID=GENE0;Name=g0")]
        with open(ann, "w") as fh:
            fh.write("##gff-version 3\n")
            for r in rows:
                fh.write("\t".join(str(x) for x in r) + "\n")
    return contigs, fa, ann


def test_genome_build_and_reload(tmp_path, capsys):
    contigs, fa, ann = _write_genome(tmp_path)
    g = guido.Genome("toy", genome_file_abspath=str(fa), annotation_file_abspath=str(ann))
    with pytest.raises(AttributeError):
        g.is_built  # upstream quirk: fai_file is only set by build()
    g.build()
    assert g.is_built and g.bowtie_index and os.path.exists(g.bowtie_index)
    assert os.path.exists(str(fa) + ".fai") and os.path.exists(str(ann) + ".gz")
    fai = [ln.split("\t") for ln in open(str(fa) + ".fai").read().splitlines()]
    assert [(r[0], int(r[1]), int(r[3]), int(r[4])) for r in fai] == [("chr1", 4000, 60, 61), ("chr2", 2500, 60, 61)]
    state = pickle.load(open(tmp_path / "toy.guido", "rb"))
    assert {"genome_name", "_bowtie_ignore", "bowtie_index", "genome_file_abspath", "annotation_file_abspath",
            "annotation_sorted_gz_file", "tbi_file", "fai_file"} <= set(state)
    g2 = guido.load_genome_from_file(tmp_path / "toy.guido")
    assert g2.__dict__ == g.__dict__
    with pytest.raises(ValueError):
        guido.load_genome_from_file(tmp_path / "missing.guido")
    with pytest.raises(ValueError):
        guido.Genome("x", genome_file_abspath=str(tmp_path / "nope.fa"))
    # the image holds the same bases as the FASTA
    img = _native.Image.load(g.bowtie_index)
    assert [(c[0], c[1]) for c in img.contigs] == [(n, len(s)) for n, s in contigs]
    assert img.fetch(0, 0, 4000) == contigs[0][1].upper() and img.fetch(1, 0, 2500) == contigs[1][1].upper()
    # sequence fetch: 1-based inclusive, case preserved
    assert g.sequence.get_seq("chr1", 11, 130).seq == contigs[0][1][10:130]


def test_locus_factories(tmp_path, cpu_fakes):
    contigs, fa, ann = _write_genome(tmp_path)
    g = guido.Genome("toy", genome_file_abspath=str(fa), annotation_file_abspath=str(ann))
    g.build()
    loc = guido.locus_from_coordinates(g, "chr1", 1100, 1600)
    assert (loc.chromosome, loc.start, loc.end) == ("chr1", 1100, 1600)
    assert loc.sequence.seq == contigs[0][1][1099:1600]
    assert set(loc.annotation["Feature"]) == {"gene", "exon"} and "Exon" in loc.annotation.columns
    assert loc.annotation["Start"].min() >= 1100 and loc.annotation["End"].max() <= 1600
    loc.find_guides(selected_features="exon")
    assert loc.intervals == [[1100, 1200], [1499, 1600]]
    loc2 = guido.locus_from_gene(g, "GENE1")
    assert (loc2.chromosome, loc2.start, loc2.end) == ("chr1", 1001, 2000)
    assert loc2.sequence.seq == contigs[0][1][1000:2000] and len(loc2.annotation) == 3
    with pytest.raises(ValueError, match="not found"):
        guido.locus_from_gene(g, "NOPE")
    assert guido.locus_from_sequence("ACGT", "s").chromosome == "s"


def test_annotate_intervals_matches_brute_force(tmp_path):
    """hits labelled with the annotation they fall into; the vectorised lookup
    against a plain per-hit filter over a GTF written to disk and read back (0-based Start)."""
    from guido_b200.annotation import annotate_intervals, read_annotation, renamed
    rng = np.random.default_rng(3)
    lines = []
    for c in ("chr1", "chr2"):
        for g in range(30):
            s = int(rng.integers(1, 60000))
            e = s + int(rng.integers(500, 4000))
            lines.append(f'{c}\tsyn\tgene\t{s}\t{e}\t.\t+\t.\tgene_id "{c}g{g}";')
            for x in range(3):
                a = s + int(rng.integers(0, e - s - 50))
                b = min(e, a + int(rng.integers(30, 300)))
                lines.append(f'{c}\tsyn\texon\t{a}\t{b}\t.\t+\t.\tgene_id "{c}g{g}"; exon_number "{x + 1}";')
                if x == 0:
                    lines.append(f'{c}\tsyn\tCDS\t{a + 3}\t{b - 3}\t.\t+\t0\tgene_id "{c}g{g}";')
    path = tmp_path / "a.gtf"
    path.write_text("\n".join(lines) + "\n")
    for df in (read_annotation(path), renamed(read_annotation(path), path)):
        idc = "ID" if "ID" in df.columns else "gene_id"
        n = 500
        ch = rng.choice(["chr1", "chr2", "chrX"], n)
        st = rng.integers(0, 66000, n)
        en = st + 23
        out = annotate_intervals(df, ch, st, en)
        assert len(out) == n
        for i in range(n):
            want = ("intergenic", "")
            for ft in ("CDS", "exon", "transcript", "gene"):
                sub = df[(df.Chromosome == ch[i]) & (df.Feature == ft) & (df.Start < en[i]) & (df.End > st[i])]
                if len(sub):
                    want = (ft, sub.sort_values("Start", kind="stable").iloc[0][idc])
                    break
            assert (out.feature[i], out.id[i]) == want, i
        assert set(out.feature) == {"intergenic", "gene", "exon", "CDS"}
    # edges: touching intervals do not overlap (End exclusive), empty inputs, empty annotation
    df = read_annotation(path)
    g0 = df[df.Feature == "gene"].iloc[0]
    o = annotate_intervals(df, [g0.Chromosome] * 2, [g0.End, g0.Start - 23], [g0.End + 23, g0.Start], features=("gene",))
    others = df[(df.Feature == "gene") & (df.Chromosome == g0.Chromosome)]
    for k, (a, b) in enumerate([(g0.End, g0.End + 23), (g0.Start - 23, g0.Start)]):
        expect = "gene" if ((others.Start < b) & (others.End > a)).any() else "intergenic"
        assert o.feature[k] == expect
    assert len(annotate_intervals(df, [], [], [])) == 0
    assert list(annotate_intervals(df.iloc[:0], ["chr1"], [5], [28]).feature) == ["intergenic"]


def test_image_load_rejects_corrupt_headers(tmp_path):
    """a .g2b whose header counts do not fit the file is refused with an error, not an abort"""
    from guido_b200 import _native as N
    img = N.Image.from_seqs([("c1", "ACGT" * 500 + "NNNN" + "GATTACA" * 30)])
    p = tmp_path / "g.g2b"
    img.save(str(p))
    raw = bytearray(p.read_bytes())
    for word, value in ((2, 1 << 60), (3, 1 << 40), (4, 1 << 50), (5, 1 << 33)):
        bad = bytearray(raw)
        bad[8 * word:8 * word + 8] = int(value).to_bytes(8, "little")
        q = tmp_path / f"bad{word}.g2b"
        q.write_bytes(bytes(bad))
        with pytest.raises(Exception):
            N.Image.load(str(q))
    (tmp_path / "short.g2b").write_bytes(bytes(raw[:len(raw) // 2]))
    with pytest.raises(Exception):
        N.Image.load(str(tmp_path / "short.g2b"))
    assert N.Image.load(str(p)).fetch(0, 0, 8) == "ACGTACGT"


def test_fasta_stray_carriage_returns(tmp_path):
    """CR bytes are skipped wherever they stand (Windows line ends, a stray CR inside a line): the contig
    length and the packed bases agree"""
    from guido_b200 import _native as N
    fa = tmp_path / "cr.fa"
    fa.write_bytes(b">c1 desc\r\nACGT\rACGT\r\nGGCC\r\n>c2\nTTTT\r\n")
    img = N.Image.from_fasta(str(fa))
    assert [(n, l) for n, l, _ in img.contigs] == [("c1", 12), ("c2", 4)]
    assert img.fetch(0, 0, 12) == "ACGTACGTGGCC" and img.fetch(1, 0, 4) == "TTTT"


def test_top_patterns_vectorised_equals_sorted(monkeypatch):
    """top_patterns_batch (all windows scored at once, winners by rounds of first-maximum per segment) ==
    sorted(generate_mmej_patterns(...), key=-score)[:n] per window: ties keep candidate order, windows with
    fewer than n candidates, windows without any"""
    from guido_b200 import mmej
    monkeypatch.setattr(_native, "mmej_batch", fakes.fake_mmej_batch)
    rng = np.random.default_rng(12)
    windows = [util.random_dna(rng, int(n)) for n in rng.integers(8, 160, size=40)]
    windows += ["ACGT" * 20, "AAAAAAAAAACCCCCCCCCC", "AC", "", "ACGTTGCA" * 10]
    cuts = [len(w) // 2 for w in windows]
    cuts[3] = 5
    for n_top in (1, 5, 1000):
        got = mmej.top_patterns_batch(cuts, windows, 20, n_top)
        for w, c, g in zip(windows, cuts, got):
            full = mmej.generate_mmej_patterns(c, w, 20)
            if full is None:
                assert g is None
            else:
                assert g == sorted(full, key=lambda x: -x["pattern_score"])[:n_top]


def test_bench_exon_window_filter_equals_the_per_site_definition():
    """bench.py's configs[4] site filter (index ranges per exon, run test on the survivors) keeps exactly
    the sites the per-site rule keeps, in the same order"""
    import bench
    rng = np.random.default_rng(14)
    P = 4
    for trial in range(12):
        fw = np.sort(rng.choice(300000, size=7000, replace=False)).astype(np.uint32) + 100
        rv = np.sort(rng.choice(300000, size=7000, replace=False)).astype(np.uint32) + 100
        ex_lo = np.sort(rng.integers(500, 299000, size=400)).astype(np.int64)
        ex_hi = ex_lo + rng.integers(50, 900, size=400)
        st = np.sort(rng.choice(300000, size=15, replace=False)).astype(np.int64)
        en = st + rng.integers(1, 400, size=15)
        keep = [0]
        for i in range(1, len(st)):
            if st[i] > en[keep[-1]]:
                keep.append(i)
        st, en = st[keep], en[keep]
        ex_end = np.minimum(ex_hi, np.append(ex_lo[1:], np.iinfo(np.int64).max))
        got = bench.window_starts_in_exons(fw, rv, P, ex_lo, ex_end, st - 75, en + 75)
        want = []
        for cut in np.concatenate([fw.astype(np.int64) - 3, rv.astype(np.int64) + P + 3]).tolist():
            k = int(np.searchsorted(ex_lo, cut, side="right")) - 1
            if k < 0 or cut >= ex_hi[k]:
                continue
            lo = cut - 75
            if any(e > lo and s < lo + 150 for s, e in zip(st.tolist(), en.tolist())):
                continue
            want.append(lo)
        assert got.tolist() == want
