"""The CPU oracle against the golden fixtures generated from the reference (tests/golden/), and
against the reference's own stored outputs.  No GPU, no /root/reference needed."""
import math

import pytest

from tests import fakes

GUIDE_ATTRS = ["id", "guide_seq", "guide_long_seq", "guide_pam", "guide_chrom", "guide_start", "guide_end",
               "guide_strand", "relative_cut_pos", "absolute_cut_pos", "locus_seq"]


def _same_guides(mine, golden, attrs=GUIDE_ATTRS):
    assert len(mine) == len(golden)
    for a, b in zip(mine, golden):
        for k in attrs:
            if k in a:
                assert a[k] == b[k], (k, a[k], b[k])


def test_kat_loci(oracle):
    kat = fakes.load_golden("kat.json")
    for key in ("locus101", "locus126", "locus490", "agap005958"):
        c = kat[key]
        mine = oracle.find_guides_records(c["seq"], c["start"], c.get("end"), chromosome=c["name"])
        _same_guides(mine, c["guides"])
    got = [[g["id"], g["guide_seq"], g["guide_start"], g["guide_end"], g["guide_strand"]]
           for g in oracle.find_guides_records(kat["locus126"]["seq"], 48714541, 48714666, chromosome="AgamP4_2R")]
    assert got == kat["locus126"]["readme_expected"]  # README.md:69-73, incl. the tie order of gRNA-1/2


def test_kat_mmej_stored_notebook_output(oracle):
    kat = fakes.load_golden("kat.json")["agap005958"]
    guides = oracle.find_guides_records(kat["seq"], kat["start"], kat["end"], chromosome=kat["name"])
    assert len(guides) == 195
    top, s, sum_score, oof = oracle.simulate_end_joining(guides[5]["locus_seq"], guides[5]["relative_cut_pos"])
    assert top == kat["stored_mmej_guide5"]           # notebooks/examples.ipynb:237-298
    assert top[3]["pattern_score"] == 230.79999999999998
    for g, m in zip(guides, kat["mmej"]):
        top, s, sum_score, oof = oracle.simulate_end_joining(g["locus_seq"], g["relative_cut_pos"])
        assert top == m["patterns"] and s == m["mmej_str"]
        assert sum_score == m["layers"]["mmej_sum_score"] and oof == m["layers"]["mmej_oof_score"]


def test_kat_cfd_record(oracle):
    from guido_b200.helpers import load_cfd_scoring_matrix
    c = fakes.load_golden("kat.json")["cfd_record"]
    mm, pam = load_cfd_scoring_matrix()
    assert oracle.cfd_score(c["guide_seq"], c["guide_pam"], c["record"]["seq"], mm, pam) == c["record"]["cfd_score"]


def test_synthetic_find_guides(oracle):
    for c in fakes.load_golden("guides_synth.json.gz"):
        mine = oracle.find_guides_records(c["seq"], c["start"], None, pam=c["pam"], chromosome="chrS",
                                          min_flanking_length=c["min_flanking_length"], intervals=c["intervals"])
        _same_guides(mine, c["guides"])
        raw = oracle.guide_records(c["seq"][:c["interval_prefix"]], c["start"], c["pam"], chromosome="chrS")
        _same_guides(raw, c["interval_guides"], [a for a in GUIDE_ATTRS if a != "id"])


def test_synthetic_mmej(oracle):
    for c in fakes.load_golden("mmej_synth.json.gz"):
        assert oracle.mmej_patterns(c["cut"], c["seq"], c["length_weight"]) == c["patterns"], (c["seq"], c["cut"])


def test_synthetic_offtargets(oracle):
    from guido_b200.helpers import load_cfd_scoring_matrix
    mm, pam_scores = load_cfd_scoring_matrix()
    for c in fakes.load_golden("offtargets_synth.json.gz"):
        contigs = [tuple(x) for x in c["contigs"]]
        for spec, golden in ((c["locus"], c["result"]), (c["external"], c["external"]["result"])):
            guides = oracle.find_guides_records(spec["seq"], spec["start"], None, pam=c["pam"], chromosome=spec["name"])
            mine = oracle.off_target_records(guides, contigs, pam=c["pam"], **c["kwargs"])
            golden = fakes.int_keys(golden)
            assert set(mine) == set(golden)
            for ix in mine:
                want = sorted(golden[ix], key=lambda r: (r["chromosome"], r["start"], r["strand"]))
                assert len(mine[ix]) == len(want)
                for a, b in zip(mine[ix], want):
                    b = dict(b, mismatches=fakes.int_keys(b["mismatches"]))
                    cfd = b.pop("cfd_score")
                    assert a == b
                    assert oracle.cfd_score(guides[ix]["guide_seq"], guides[ix]["guide_pam"], a["seq"], mm, pam_scores) == cfd


def test_bowtie_tsv_format_matches_stored_rows(oracle):
    """Format of a '-' strand row with mismatches, as stored in the reference's test.ipynb:14081:
    read TTATCATCCACTCTGACGGGTGG, shown reverse-complemented, descriptors on the forward strand."""
    read = "TTATCATCCACTCTGACGGGTGG"
    # build a reference window that differs from rc(read) at the stored offsets 15,17,18,19,20,22
    shown = oracle.rev_comp(read)
    assert shown == "CCACCCGTCAGAGTGGATGATAA"
    stored = {15: ("G", "T"), 17: ("T", "C"), 18: ("T", "C"), 19: ("G", "C"), 20: ("C", "A"), 22: ("T", "C")}
    window = list(shown)
    for off, (ref, rd) in stored.items():
        assert shown[22 - off] == rd
        window[22 - off] = ref
    genome = "ACGT" * 10 + "".join(window) + "TGCA" * 10
    raw = oracle.bowtie_n([("2R", genome)], [read], 1, 13, 180)
    rows = oracle.bowtie_tsv(raw, ["0"], [read]).strip().split("\n")
    assert "0\t-\t2R\t40\tCCACCCGTCAGAGTGGATGATAA\t" + "I" * 23 + "\t0\t15:G>T,17:T>C,18:T>C,19:G>C,20:C>A,22:T>C" in rows


def test_bowtie_rule_budget(oracle):
    """-e 150 admits 5 mismatches incl. the read's N, not 6; -n counts N inside the seed."""
    guide = "ACGTTGCATGCAAGCTTGCA"
    read = oracle.rev_comp(guide + "NGG")
    def genome(mm_positions):
        g = list(guide)
        for p in mm_positions:
            g[p] = "ACGT"[("ACGT".index(g[p]) + 1) % 4]
        return "TTTT" * 8 + "".join(g) + "AGG" + "TTTT" * 8
    args = oracle.bowtie_args("NGG", 10, 2, 4)
    assert args == (3, 13, 150)
    assert len(oracle.bowtie_n([("c", genome([0, 1, 2, 3]))], [read], *args)) == 1      # 4 distal
    assert len(oracle.bowtie_n([("c", genome([0, 1, 2, 3, 4]))], [read], *args)) == 0   # 5 distal > e
    assert len(oracle.bowtie_n([("c", genome([18, 19]))], [read], *args)) == 1          # 2 in core
    assert len(oracle.bowtie_n([("c", genome([17, 18, 19]))], [read], *args)) == 0      # 3 in core > n
    assert len(oracle.bowtie_n([("c", genome([9, 18, 19]))], [read], *args)) == 1       # position 10 is distal
    with pytest.raises(ValueError):
        oracle.bowtie_args("NGG", 10, 3, 4)
    with pytest.raises(ValueError):
        oracle.bowtie_args("NGG", 10, 2, 1)


def test_seed_index_port_equals_brute_force(oracle):
    """the seed-indexed CPU search (bench.py's CPU arm) reports exactly the alignments of the
    brute-force oracle that survive guido's PAM filter -- both strands, N runs, IUPAC letters,
    several parameter sets"""
    import numpy as np
    from tests import util
    contigs = util.synthetic_contigs(71, [90000, 30000, 400], n_runs=4, iupac=5)
    rng = np.random.default_rng(5)
    protos = []
    s0 = contigs[0][1].upper()
    while len(protos) < 150:
        p = int(rng.integers(0, len(s0) - 23))
        w = s0[p:p + 20]
        if set(w) - set("ACGT"):
            continue
        w = list(w)
        for _ in range(int(rng.integers(0, 5))):
            w[int(rng.integers(0, 20))] = "ACGT"[int(rng.integers(0, 4))]
        protos.append("".join(w))
    P = np.frombuffer("".join(protos).encode(), np.uint8).reshape(-1, 20)
    for pam, kw in (("NGG", {}), ("NAG", dict(core_mismatches=1, total_mismatches=3)), ("TGG", dict(core_mismatches=3))):
        valid, _ = util.oracle_valid_hits(oracle, contigs, protos, pam, **kw)
        ix = oracle.SeedIndex([(n, s, i << 24) for i, (n, s) in enumerate(contigs)], pam, kw.get("core_length", 10))
        n_unk = pam.count("N")
        hits = ix.search(P, kw.get("core_mismatches", 2), kw.get("total_mismatches", 4) + 1 - n_unk, want_hits=True)
        assert ix.search(P, kw.get("core_mismatches", 2), kw.get("total_mismatches", 4) + 1 - n_unk) == len(hits)
        got = {(int(h >> 33), contigs[int((h >> 1) & 0xffffffff) >> 24][0], int((h >> 1) & 0xffffff), "-" if h & 1 else "+")
               for h in hits.tolist()}
        assert got == valid and len(valid) > 0


def test_bowtie_rows_fixture(oracle):
    """raw rows of a REAL bowtie 1.3.1 run (written by tools/compare_with_bowtie.py --write-golden on a
    machine that has bowtie) pin oracle.bowtie_n; skipped while no fixture has been generated"""
    import glob
    import gzip
    import json
    import os
    import pytest
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "bowtie_rows", "*.json.gz")))
    if not files:
        pytest.skip("no bowtie fixture yet: run tools/compare_with_bowtie.py --write-golden where bowtie 1.3.1 is installed")
    for f in files:
        g = json.load(gzip.open(f, "rt"))
        contigs = [tuple(c) for c in g["contigs"]]
        for case in g["cases"]:
            pam = case["pam"]
            seed_mms, seed_len, e = oracle.bowtie_args(pam, case["core_length"], case["core_mismatches"], case["total_mismatches"])
            reads = oracle.reads_for_guides([p + pam for p in g["protos"]], pam)
            names = [str(i) for i in range(len(reads))]
            ours = oracle.bowtie_tsv(oracle.bowtie_n(contigs, reads, seed_mms, seed_len, e), names, reads)
            rows = lambda tsv: {tuple(l.split("\t")[i] for i in (0, 1, 2, 3, 7)) for l in tsv.splitlines() if l.count("\t") >= 7}
            assert rows(ours) == rows(case["tsv"]), (f, case["pam"])
