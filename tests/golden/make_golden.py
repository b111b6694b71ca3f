"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference (nkran/guido, pure Python) is imported through ``oracle/ref_loader.py`` (stub
modules for the absent third-party imports, no algorithmic content).  For the off-target fixture
its real ``run_bowtie`` / ``Locus.find_off_targets`` run against ``oracle/fake_bowtie.py`` (bowtie
itself is not in the image), so everything downstream of the alignment rule is reference code.
Stored outputs of the reference's own notebooks / docs are copied in as data (known answers).
"""
import ast
import gzip
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import fake_bowtie, ref_loader  # noqa: E402
from tests import util  # noqa: E402

REF = ref_loader.REFERENCE_ROOT
GUIDE_ATTRS = ["id", "guide_seq", "guide_long_seq", "guide_pam", "guide_chrom", "guide_start", "guide_end",
               "guide_strand", "relative_cut_pos", "absolute_cut_pos", "locus_seq"]


def jsonable(x):
    if isinstance(x, dict):
        return {str(k): jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.floating, float)):
        x = float(x)
        return {"__float__": x.hex()} if not math.isnan(x) else {"__float__": "nan"}
    return x


def guide_rec(g):
    return {a: getattr(g, a) for a in GUIDE_ATTRS}


def notebook_outputs(path):
    nb = json.load(open(path))
    return nb["cells"]


def main():
    guido = ref_loader.load_reference()
    import pandas as pd

    # ------------------------------------------------------------------ known answers (stored)
    kat = {}
    seq101 = ("TTATCATCCACTCTGACGGGTGGTATTGCGCAACTCCACGCCATCAAACATGTTCAGATTATGCAATCGTGAGTATTCGTTGACCACCGCTT"
              "GACCTGTGT")
    loc = guido.Locus(sequence=seq101, name="AgamP4_2R", start=48714554, end=48714654)
    kat["locus101"] = {"source": "input guido/locus.py:60-63", "seq": seq101, "name": "AgamP4_2R",
                       "start": 48714554, "end": 48714654, "guides": [guide_rec(g) for g in loc.find_guides()]}

    # 490-bp mixed-case AgamP4 string, commented out in `test API.ipynb`
    seq490 = None
    for cell in notebook_outputs(os.path.join(REF, "test API.ipynb")):
        for line in cell.get("source", []):
            if line.strip().startswith('# seq = "GCCGACCCATTCTGCTG'):
                seq490 = line.split('"')[1]
    assert seq490 and len(seq490) == 490, len(seq490 or "")
    a = seq490.index("TGGTGCGGAAAGTTTAT")
    b = seq490.index("TGTGTTAAACATAAATG") + len("TGTGTTAAACATAAATG")
    seq126 = seq490[a:b]
    assert len(seq126) == 126
    loc = guido.Locus(sequence=seq126, name="AgamP4_2R", start=48714541, end=48714666)
    kat["locus126"] = {
        "source": "sequence test API.ipynb:37; expected list README.md:69-73 / tests/test_guido.py:98",
        "seq": seq126, "name": "AgamP4_2R", "start": 48714541, "end": 48714666,
        "readme_expected": [
            ["gRNA-1", "AAGTTTATCATCCACTCTGACGG", 48714550, 48714572, "+"],
            ["gRNA-2", "CGCAATACCACCCGTCAGAGTGG", 48714561, 48714583, "-"],
            ["gRNA-3", "AGTTTATCATCCACTCTGACGGG", 48714551, 48714573, "+"],
            ["gRNA-4", "TTATCATCCACTCTGACGGGTGG", 48714554, 48714576, "+"],
            ["gRNA-5", "TCTGAACATGTTTGATGGCGTGG", 48714589, 48714611, "-"],
            ["gRNA-6", "CATAATCTGAACATGTTTGATGG", 48714594, 48714616, "-"],
            ["gRNA-7", "GTTTAACACAGGTCAAGCGGTGG", 48714637, 48714659, "-"],
            ["gRNA-8", "TATGTTTAACACAGGTCAAGCGG", 48714640, 48714662, "-"]],
        "guides": [guide_rec(g) for g in loc.find_guides()]}
    got = [[g.id, g.guide_seq, g.guide_start, g.guide_end, g.guide_strand] for g in loc.guides]
    assert got == kat["locus126"]["readme_expected"], got
    loc = guido.Locus(sequence=seq490, name="AgamP4_2R", start=48714364)
    kat["locus490"] = {"source": "test API.ipynb:37 (soft-masked lower case)", "seq": seq490, "name": "AgamP4_2R",
                       "start": 48714364, "guides": [guide_rec(g) for g in loc.find_guides()]}

    # AGAP005958 locus + stored MMEJ top-5 of guide(5) (notebooks/examples.ipynb)
    agap_seq, stored_mmej = None, None
    cells = notebook_outputs(os.path.join(REF, "notebooks", "examples.ipynb"))
    for cell in cells:
        src = "".join(cell.get("source", []))
        for out in cell.get("outputs", []):
            text = "".join(out.get("data", {}).get("text/plain", []))
            if src.strip() == "l.sequence.seq":
                agap_seq = ast.literal_eval(text)
            if src.strip() == "l.guide(5).mmej_patterns":
                stored_mmej = ast.literal_eval(text)
    assert agap_seq and len(agap_seq) == 1684 and stored_mmej and len(stored_mmej) == 5
    loc = guido.Locus(sequence=agap_seq, name="AgamP4_2L", start=24049834, end=24051517)
    loc.find_guides()
    loc.simulate_end_joining()
    assert loc.guide(5).mmej_patterns == stored_mmej
    kat["agap005958"] = {
        "source": "notebooks/examples.ipynb:192 (sequence), :237-298 (stored MMEJ top-5 of guide 5)",
        "seq": agap_seq, "name": "AgamP4_2L", "start": 24049834, "end": 24051517,
        "stored_mmej_guide5": stored_mmej,
        "guides": [guide_rec(g) for g in loc.guides],
        "mmej": [{"patterns": g.mmej_patterns, "mmej_str": g.mmej_str, "layers": dict(g.layers)} for g in loc.guides]}
    kat["cfd_record"] = {
        "source": "docs/source/getting_started.rst:104-112",
        "guide_seq": "AAGTTTATCATCCACTCTGACGG", "guide_pam": "CGG",
        "record": {"ix": 0, "mismatches": {21: "A", 20: "G", 18: "G", 10: "C", 6: "C"},
                   "mismatches_string": ".....C...C.......G.GA..", "chromosome": "AgamP4_2L", "start": 8079046,
                   "strand": "+", "seq": "AAGTTCATCCTCCACTCGGGAGG", "cfd_score": 0.07724301822162806}}
    json.dump(jsonable(kat), open(os.path.join(HERE, "kat.json"), "w"), indent=1)

    # ------------------------------------------------------------------ find_guides on synthetic loci
    cases = []
    specs = [(101, 3000, "NGG", 0, 1, None), (102, 5000, "NGG", 75, 1000, None), (103, 2500, "NAG", 10, 50, None),
             (104, 4000, "TTTV", 0, 1, None), (105, 3000, "NRG", 0, 7, None), (106, 800, "NGG", 30, 1, "intervals"),
             (107, 60, "NGG", 0, 1, None), (108, 22, "NGG", 0, 1, None), (109, 2000, "YGN", 5, 123456, None),
             (110, 6000, "NGG", 0, 1, "iupac")]
    for seed, n, pam, flank, start, extra in specs:
        contigs = util.synthetic_contigs(seed, [n], n_runs=2, run_len=(3, 40), lower_frac=0.2,
                                         iupac=8 if extra == "iupac" else 0)
        seq = contigs[0][1]
        ann = None
        feats = None
        if extra == "intervals":
            ann = pd.DataFrame({"Chromosome": ["chrS"] * 4, "Feature": ["exon", "exon", "exon", "CDS"],
                                "Start": [start + 100, start + 180, start + 500, start + 50],
                                "End": [start + 200, start + 260, start + 640, start + 700]})
            feats = "exon"
        loc = guido.Locus(sequence=seq, name="chrS", start=start, annotation=ann)
        guides = loc.find_guides(pam=pam, min_flanking_length=flank, selected_features=feats or "all")
        raw = loc._find_guides_in_interval(seq[: min(n, 400)], start, pam)
        cases.append({"seed": seed, "seq": seq, "start": start, "pam": pam, "min_flanking_length": flank,
                      "annotation": None if ann is None else ann.to_dict(orient="list"), "selected_features": feats,
                      "intervals": loc.intervals, "guides": [guide_rec(g) for g in guides],
                      "interval_prefix": min(n, 400), "interval_guides": [guide_rec(g) for g in raw]})
    with gzip.open(os.path.join(HERE, "guides_synth.json.gz"), "wt") as fh:
        json.dump(jsonable(cases), fh)

    # ------------------------------------------------------------------ MMEJ on synthetic windows
    rng = np.random.default_rng(7)
    mm_cases = []
    windows = [(util.random_dna(rng, 150), 75) for _ in range(12)]
    windows += [(util.random_dna(rng, n), c) for n, c in [(150, 60), (120, 45), (90, 75), (40, 20), (10, 5), (5, 2), (4, 3)]]
    windows += [("A" * 150, 75), ("AT" * 75, 75), ("ACG" * 50, 75), ("GATTACA" * 21, 70), ("AC" * 20 + "GT" * 20, 40)]
    for seq, cut in windows:
        for lw in (20, 13.5):
            pats = guido.mmej.generate_mmej_patterns(cut, seq, lw)
            mm_cases.append({"seq": seq, "cut": cut, "length_weight": lw, "patterns": pats})
    with gzip.open(os.path.join(HERE, "mmej_synth.json.gz"), "wt") as fh:
        json.dump(jsonable(mm_cases), fh)

    # ------------------------------------------------------------------ off-targets via fake bowtie
    ot_cases = []
    for seed, pam, kw, lens in [(201, "NGG", {}, [30000, 30000, 8000]),
                                (202, "NGG", dict(core_length=10, core_mismatches=2, total_mismatches=3), [40000, 20000]),
                                (203, "NAG", dict(core_length=12, core_mismatches=1, total_mismatches=4), [25000, 25000])]:
        contigs = util.synthetic_contigs(seed, lens, n_runs=3, run_len=(5, 100), lower_frac=0.1)
        # plant near-copies of a locus region on the other contigs so that real hits exist
        rng = np.random.default_rng(seed)
        name0, s0 = contigs[0]
        region = s0[5000:5600].upper()
        planted = []
        for ci in range(1, len(contigs)):
            nm, s = contigs[ci]
            s = list(s)
            for rep in range(3):
                pos = int(rng.integers(100, len(s) - 700))
                frag = list(region if rep % 2 == 0 else guido.rev_comp(region))
                for _ in range(20):
                    frag[int(rng.integers(0, len(frag)))] = "ACGT"[int(rng.integers(0, 4))]
                s[pos:pos + len(frag)] = frag
            planted.append((nm, "".join(s)))
        contigs = [contigs[0]] + planted
        index = f"/virtual/index/{seed}"
        loc = guido.Locus(sequence=s0[4990:5610], name=name0, start=4991, genome=fake_bowtie.FakeGenome(index))
        loc.find_guides(pam=pam)
        with fake_bowtie.patched(guido, {index: contigs}) as fake:
            result = loc.find_off_targets(**kw)
            argv = fake.last_argv
        # external-genome style search: guides from an unrelated locus (most have no key)
        other = util.synthetic_contigs(seed + 50, [700], n_runs=0)[0][1]
        loc2 = guido.Locus(sequence=other, name="elsewhere", start=1, genome=fake_bowtie.FakeGenome(index))
        loc2.find_guides(pam=pam)
        with fake_bowtie.patched(guido, {index: contigs}):
            result2 = loc2.find_off_targets(**kw)
        ot_cases.append({
            "seed": seed, "pam": pam, "kwargs": kw, "contigs": contigs, "argv": argv[:-1],
            "locus": {"seq": s0[4990:5610], "name": name0, "start": 4991},
            "result": result, "off_target_str": [g.off_target_str for g in loc.guides],
            "off_targets_string": [g.off_targets_string for g in loc.guides],
            "layers": [dict(g.layers) for g in loc.guides],
            "external": {"seq": other, "name": "elsewhere", "start": 1, "result": result2,
                         "layers": [dict(g.layers) for g in loc2.guides]}})
    with gzip.open(os.path.join(HERE, "offtargets_synth.json.gz"), "wt") as fh:
        json.dump(jsonable(ot_cases), fh)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
