"""Parity of the public API on a real B200: the shipped path (Python API -> C-ABI -> sm_100a
kernels) against the golden fixtures generated from the reference and against the CPU oracle."""
import math

import numpy as np
import pandas as pd
import pytest

import guido_b200 as guido
from guido_b200 import _native, off_targets
from tests import fakes, util
from tests.test_host_logic import GUIDE_ATTRS, _write_genome, same_guides, same_layers

pytestmark = pytest.mark.gpu


def test_kat_loci_and_mmej():
    kat = fakes.load_golden("kat.json")
    for key in ("locus101", "locus126", "locus490", "agap005958"):
        c = kat[key]
        loc = guido.Locus(sequence=c["seq"], name=c["name"], start=c["start"], end=c.get("end"))
        same_guides(loc.find_guides(), c["guides"])
    loc.simulate_end_joining()
    assert loc.guide(5).mmej_patterns == kat["agap005958"]["stored_mmej_guide5"]
    for g, m in zip(loc.guides, kat["agap005958"]["mmej"]):
        assert g.mmej_patterns == m["patterns"] and g.mmej_str == m["mmej_str"]
        same_layers(g.layers, m["layers"])


def test_synthetic_find_guides_golden():
    for c in fakes.load_golden("guides_synth.json.gz"):
        ann = pd.DataFrame(c["annotation"]) if c["annotation"] else None
        loc = guido.Locus(sequence=c["seq"], name="chrS", start=c["start"], annotation=ann)
        guides = loc.find_guides(pam=c["pam"], min_flanking_length=c["min_flanking_length"],
                                 selected_features=c["selected_features"] or "all")
        same_guides(guides, c["guides"])
        raw = loc._find_guides_in_interval(c["seq"][:c["interval_prefix"]], c["start"], c["pam"])
        same_guides(raw, c["interval_guides"], [a for a in GUIDE_ATTRS if a != "id"])


def test_synthetic_mmej_golden():
    from guido_b200.mmej import generate_mmej_patterns, generate_mmej_patterns_batch
    cases = fakes.load_golden("mmej_synth.json.gz")
    for lw in (20, 13.5):
        sub = [c for c in cases if c["length_weight"] == lw]
        got = generate_mmej_patterns_batch([c["cut"] for c in sub], [c["seq"] for c in sub], lw)
        assert got == [c["patterns"] for c in sub]
    assert generate_mmej_patterns(cases[0]["cut"], cases[0]["seq"], 20) == cases[0]["patterns"]


@pytest.mark.parametrize("algo", [_native.OT_SCAN, _native.OT_AUTO])
def test_offtargets_golden(tmp_path, algo):
    for c in fakes.load_golden("offtargets_synth.json.gz"):
        contigs = [tuple(x) for x in c["contigs"]]
        path = tmp_path / f"g{c['seed']}.g2b"
        _native.Image.from_seqs(contigs).save(path)
        genome = guido.Genome("g", bowtie_index_abspath=str(path))
        for spec, golden, layers in ((c["locus"], c["result"], c["layers"]),
                                     (c["external"], c["external"]["result"], c["external"]["layers"])):
            loc = guido.Locus(sequence=spec["seq"], name=spec["name"], start=spec["start"], genome=genome)
            loc.find_guides(pam=c["pam"])
            found = loc.find_off_targets(algo=algo, **c["kwargs"])
            golden = fakes.int_keys(golden)
            assert set(found) == set(golden)
            key = lambda r: (r["chromosome"], r["start"], r["strand"])
            for ix, recs in golden.items():
                want = sorted((dict(r, mismatches=fakes.int_keys(r["mismatches"])) for r in recs), key=key)
                assert sorted(found[ix], key=key) == want
            for g, lay in zip(loc.guides, layers):
                if "ot_cfd_score_mean" in lay and not math.isnan(lay["ot_cfd_score_mean"]):
                    assert g.layers["ot_sum_score"] == lay["ot_sum_score"]
                    assert g.layers["ot_cfd_score_max"] == lay["ot_cfd_score_max"]
                    assert math.isclose(g.layers["ot_cfd_score_sum"], lay["ot_cfd_score_sum"], rel_tol=1e-12)
                else:
                    same_layers(g.layers, lay)
            if spec is c["locus"]:
                assert [g.off_target_str for g in loc.guides] == c["off_target_str"]
                assert [g.off_targets_string for g in loc.guides] == c["off_targets_string"]


def test_end_to_end_genome_build_to_offtargets(tmp_path, oracle):
    """Genome.build -> locus_from_gene -> find_guides -> find_off_targets -> simulate_end_joining
    on a toy FASTA/GFF3, every record checked against the oracle."""
    contigs, fa, ann = _write_genome(tmp_path)
    g = guido.Genome("toy", genome_file_abspath=str(fa), annotation_file_abspath=str(ann))
    g.build()
    loc = guido.locus_from_gene(g, "GENE1")
    guides = loc.find_guides(selected_features="exon", min_flanking_length=20)
    want = oracle.find_guides_records(loc.sequence.seq, loc.start, loc.end, chromosome="chr1",
                                      min_flanking_length=20, intervals=loc.intervals)
    assert [(x.id, x.guide_seq, x.guide_start, x.guide_end, x.guide_strand) for x in guides] == \
        [(w["id"], w["guide_seq"], w["guide_start"], w["guide_end"], w["guide_strand"]) for w in want]
    found = loc.find_off_targets(total_mismatches=4)
    exp = oracle.off_target_records(want, contigs, total_mismatches=4)
    assert set(found) == set(exp)
    for ix in exp:
        got = [{k: v for k, v in r.items() if k != "cfd_score"} for r in found[ix]]
        assert sorted(got, key=lambda r: (r["chromosome"], r["start"], r["strand"])) == exp[ix]
    loc.simulate_end_joining()
    for x, w in zip(loc.guides, want):
        top = oracle.simulate_end_joining(w["locus_seq"], w["relative_cut_pos"])
        assert x.mmej_patterns == top[0] and x.layers["mmej_sum_score"] == top[2]


def test_bad_arguments_raise_like_the_reference(tmp_path):
    contigs = util.synthetic_contigs(9, [3000], n_runs=0)
    path = tmp_path / "g.g2b"
    _native.Image.from_seqs(contigs).save(path)
    genome = guido.Genome("g", bowtie_index_abspath=str(path))
    loc = guido.Locus(sequence=contigs[0][1][:300], name="other", genome=genome)
    loc.find_guides()
    with pytest.raises(ValueError, match="core_mismatches is not valid: 3"):
        loc.find_off_targets(core_mismatches=3)
    with pytest.raises(ValueError, match="cannot be greater than total_mismatches: 2 > 1"):
        loc.find_off_targets(total_mismatches=1)
    with pytest.raises(KeyError):
        loc.find_guides(pam="NGZ")
    # a PAM without N: the self-hit is a perfect alignment and upstream crashes on it
    loc2 = guido.Locus(sequence=contigs[0][1][:300], name="other", genome=genome)
    if loc2.find_guides(pam="AGG"):
        with pytest.raises(ValueError, match="not enough values"):
            loc2.find_off_targets()
