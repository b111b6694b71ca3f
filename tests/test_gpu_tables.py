"""Columnar tables (guido_b200/tables.py) against the per-locus API: same guides, same records."""
import numpy as np
import pytest

import guido_b200 as guido
from guido_b200 import _native, off_targets, tables
from tests import util
from tests.test_host_logic import GUIDE_ATTRS

pytestmark = pytest.mark.gpu


def test_guide_table_equals_find_guides_per_contig(tmp_path):
    contigs = util.synthetic_contigs(61, [30000, 12000, 50, 8000], n_runs=3, run_len=(5, 300), lower_frac=0.2)
    img = _native.Image.from_seqs(contigs).upload()
    t = tables.GuideTable(img, "NGG")
    k = 0
    for ci, (name, seq) in enumerate(contigs):
        loc = guido.Locus(sequence=seq, name=name)
        want = loc.find_guides()
        n = int((t.contig == ci).sum())
        assert n == len(want)
        for j, w in enumerate(want):
            g = t[k + j]
            for a in GUIDE_ATTRS:
                assert getattr(g, a) == getattr(w, a), (name, j, a, getattr(g, a), getattr(w, a))
        protos = t.protospacers(np.arange(k, k + n))
        assert [p.tobytes().decode() for p in protos] == [w.guide_seq[:20] for w in want]
        k += n
    assert k == len(t)
    # exporters straight from the columns: same rows as the per-locus exporters, contig by contig
    df = t.to_dataframe(with_sequences=True)
    bed = tmp_path / "all.bed"
    t.to_bed(bed)
    bed_lines = bed.read_text().splitlines()
    k = 0
    for ci, (name, seq) in enumerate(contigs):
        loc = guido.Locus(sequence=seq, name=name)
        loc.find_guides()
        ref = loc.guides_to_dataframe()
        n = len(ref)
        mine = df.iloc[k:k + n]
        for col in ("guide_chrom", "guide_start", "guide_end", "guide_strand", "absolute_cut_pos", "id", "guide_seq"):
            assert mine[col].tolist() == (ref[col].tolist() if n else []), (name, col)
        if n:  # (an empty guide list has no columns to pick: guides_to_bed raises, as upstream)
            loc.guides_to_bed(tmp_path / "one.bed")
            assert bed_lines[k:k + n] == (tmp_path / "one.bed").read_text().splitlines()
        k += n
    assert len(tables.GuideTable(img, "TTTV")) == 0  # upstream: only 3-nt PAMs give guides
    with pytest.raises(KeyError):
        tables.GuideTable(img, "NGX")


def test_offtarget_table_equals_run_bowtie(tmp_path):
    contigs = util.synthetic_contigs(62, [60000, 40000], n_runs=2)
    s0 = contigs[0][1].upper()
    contigs[1] = (contigs[1][0], contigs[1][1][:5000] + s0[1000:1800] + contigs[1][1][5800:])
    path = tmp_path / "g.g2b"
    _native.Image.from_seqs(contigs).save(path)
    genome = guido.Genome("g", bowtie_index_abspath=str(path))
    loc = guido.Locus(sequence=contigs[0][1][990:1810], name="chr1", start=991, genome=genome)
    guides = loc.find_guides()
    found = loc.find_off_targets()
    img = off_targets.get_image(str(path))
    t = tables.offtarget_table(img, [g.guide_seq[:20] for g in guides], guide_chrom=[g.guide_chrom for g in guides],
                               guide_start=[g.guide_start for g in guides])
    assert set(np.nonzero(t.has_key)[0].tolist()) == set(found)
    sums = t.ot_sum_score()
    counts = t.counts_matrix()
    key = lambda r: (r["chromosome"], r["start"], r["strand"])
    total = 0
    for ix, recs in found.items():
        want = sorted(({k: v for k, v in r.items() if k != "cfd_score"} for r in recs), key=key)
        assert sorted(t.records(ix), key=key) == want
        assert sums[ix] == guides[ix].layers["ot_sum_score"]
        assert "|".join(str(c) for c in counts[ix]) == guides[ix].off_target_str
        total += len(recs)
    assert total == len(t) and total > 50
    # vectorised CFD: every per-hit score bit-identical to the per-guide loop, aggregates to 1e-12
    cfd = t.cfd_scores(guide_pams=[g.guide_pam for g in guides])
    mean, best, tot = t.cfd_layers(guide_pams=[g.guide_pam for g in guides])
    for ix, recs in found.items():
        mine = sorted(zip((key(r) for r in t.records(ix)), cfd[t.guide == ix].tolist()))
        want = sorted((key(r), r["cfd_score"]) for r in recs)
        assert mine == want
        if recs:
            assert best[ix] == guides[ix].layers["ot_cfd_score_max"]
            assert np.isclose(mean[ix], guides[ix].layers["ot_cfd_score_mean"], rtol=1e-12, atol=0)
            assert np.isclose(tot[ix], guides[ix].layers["ot_cfd_score_sum"], rtol=1e-12, atol=0)
        else:
            assert np.isnan(mean[ix]) and np.isnan(best[ix])
    # hits against an annotation: the planted copy on chr2 lies in a "gene"
    import pandas as pd
    c2 = contigs[1][0]
    ann = pd.DataFrame([dict(Chromosome=c2, Feature="gene", Start=4990, End=5810, ID="planted"),
                        dict(Chromosome=c2, Feature="exon", Start=5100, End=5200, ID="planted")])
    lab = t.annotate(ann)
    assert len(lab) == len(t) and list(lab.columns) == ["guide", "chromosome", "start", "strand", "feature", "id"]
    on2 = (t.chromosome == c2)
    inside = on2 & (t.start < 5810) & (t.start + 23 > 4990)
    in_exon = on2 & (t.start < 5200) & (t.start + 23 > 5100)
    assert inside.sum() > 20 and in_exon.sum() > 0
    assert np.array_equal(lab.feature.to_numpy() == "exon", in_exon)
    assert np.array_equal(lab.feature.to_numpy() == "gene", inside & ~in_exon)
    assert np.array_equal(lab.feature.to_numpy() == "intergenic", ~inside)
    assert set(lab.id[inside]) == {"planted"} and set(lab.id[~inside]) <= {""}
