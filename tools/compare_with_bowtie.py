#!/usr/bin/env python
"""compare_with_bowtie.py -- pin the bowtie alignment rule against a REAL bowtie 1.3.1.

The one step of the hot path whose parity is unpinned in the build container is the alignment rule
itself: bowtie 1.3.1 is an external binary the reference shells out to
(/root/reference/guido/off_targets.py:151-167) and it is not in the image.  Given a `bowtie` and
`bowtie-build` on PATH (or --bowtie-path), this script

  1. writes a seeded synthetic genome (N runs, IUPAC letters, lower case) and guides that sit on it
     with 0..5 substitutions -- including guides with mismatches at the seed boundary, N in the PAM,
     total+1 mismatches, windows that touch N runs and contig ends,
  2. builds a bowtie index and runs the EXACT command line the reference builds
     (`bowtie -p T --quiet -y -n S -l L -e E -a -x IDX -f reads.fa`, off_targets.py:151-158),
  3. parses the raw rows and compares, row for row, with
        * oracle.bowtie_n (the C restatement the tests use as the checker), and
        * the GPU search (guido_b200, if a device is present): valid hits + key presence flags,
  4. with --write-golden stores the raw TSV and the inputs under tests/golden/bowtie_rows/ so the
     CPU test-suite can pin oracle.bowtie_n against them from then on
     (tests/test_oracle_golden.py::test_bowtie_rows_fixture picks the files up when they exist).

Exit code 0 = every compared set is identical.  Nothing here runs in CI of this repo (no bowtie);
it is what lets SURVEY 8(a5)/(c) go from "parity unpinned" to pinned on a machine that has bowtie.
"""
import argparse
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [  # (pam, core_length, core_mismatches, total_mismatches)
    ("NGG", 10, 2, 4), ("NGG", 10, 2, 3), ("NGG", 12, 1, 5), ("NGG", 10, 0, 0),
    ("NAG", 10, 2, 4), ("NRG", 10, 1, 4), ("TGG", 10, 3, 4),
]


def make_inputs(seed, genome_kb, n_guides):
    from tests import util
    lens = [genome_kb * 600, genome_kb * 300, genome_kb * 100, 700]  # bases
    contigs = util.synthetic_contigs(seed, lens, n_runs=6, run_len=(5, 400), lower_frac=0.1, iupac=8)
    rng = np.random.default_rng(seed + 1)
    protos = []
    while len(protos) < n_guides:
        ci = int(rng.integers(0, len(contigs)))
        s = contigs[ci][1].upper()
        p = int(rng.integers(0, max(len(s) - 23, 1)))
        w = s[p:p + 20]
        if len(w) < 20 or set(w) - set("ACGT"):
            continue
        w = list(w if rng.random() < 0.5 else w[::-1].translate(str.maketrans("ACGT", "TGCA")))
        k = int(rng.integers(0, 6))
        # mismatches biased towards the seed boundary (guide positions 10/11) and the distal end
        pos = rng.choice([0, 1, 8, 9, 10, 11, 12, 18, 19] + list(range(20)), size=k, replace=False) if k else []
        for q in pos:
            w[int(q)] = "ACGT"[int(rng.integers(0, 4))]
        protos.append("".join(w))
    return contigs, protos


def run_bowtie(bowtie_path, contigs, reads, names, seed_mms, seed_len, e, threads, workdir):
    fa = os.path.join(workdir, "genome.fa")
    with open(fa, "w") as fh:
        for name, s in contigs:
            fh.write(f">{name}\n")
            fh.writelines(s[i:i + 60] + "\n" for i in range(0, len(s), 60))
    idx = os.path.join(workdir, "genome")
    bb = os.path.join(bowtie_path, "bowtie-build") if bowtie_path else "bowtie-build"
    bt = os.path.join(bowtie_path, "bowtie") if bowtie_path else "bowtie"
    subprocess.run([bb, "--quiet", fa, idx], check=True, stdout=subprocess.DEVNULL)
    rf = os.path.join(workdir, "reads.fa")
    with open(rf, "w") as fh:
        for n, r in zip(names, reads):
            fh.write(f">{n}\n{r}\n")
    # the reference's argv, verbatim (off_targets.py:151-158)
    argv = [bt, "-p", str(threads), "--quiet", "-y", "-n", str(seed_mms), "-l", str(seed_len), "-e", str(e),
            "-a", "-x", idx, "-f", rf]
    out = subprocess.run(argv, check=True, capture_output=True, text=True).stdout
    return out, argv


def parse_rows(tsv):
    rows = set()
    for line in tsv.splitlines():
        f = line.split("\t")
        if len(f) < 7:
            continue
        rows.add((f[0], f[1], f[2], int(f[3]), f[7] if len(f) > 7 else ""))
    return rows


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--bowtie-path", default="", help="directory holding bowtie and bowtie-build (default: PATH)")
    ap.add_argument("--seed", type=int, default=4242)
    ap.add_argument("--genome-kb", type=int, default=1000, help="genome size in kb (default 1 Mb in four contigs)")
    ap.add_argument("--guides", type=int, default=400)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--write-golden", action="store_true")
    args = ap.parse_args()
    bt = os.path.join(args.bowtie_path, "bowtie") if args.bowtie_path else shutil.which("bowtie")
    if not bt or not os.path.exists(bt):
        print("bowtie not found (PATH / --bowtie-path): nothing compared.  Install bowtie 1.3.1 and run again.")
        return 2
    from oracle import oracle as O
    contigs, protos = make_inputs(args.seed, args.genome_kb, args.guides)
    gpu = None
    try:
        from guido_b200 import _native as N
        if N.device_count() > 0:
            gpu = N.Image.from_seqs(contigs).upload()
    except Exception as exc:  # the comparison with the oracle still runs
        print(f"(GPU comparison skipped: {exc})")
    bad = 0
    golden = {"seed": args.seed, "contigs": contigs, "protos": protos, "cases": []}
    with tempfile.TemporaryDirectory() as wd:
        for pam, cl, cm, tm in CASES:
            seed_mms, seed_len, e = O.bowtie_args(pam, cl, cm, tm)
            reads = O.reads_for_guides([p + pam for p in protos], pam)
            names = [str(i) for i in range(len(reads))]
            tsv, argv = run_bowtie(args.bowtie_path, contigs, reads, names, seed_mms, seed_len, e, args.threads, wd)
            real = parse_rows(tsv)
            ours = parse_rows(O.bowtie_tsv(O.bowtie_n(contigs, reads, seed_mms, seed_len, e), names, reads))
            tag = f"{pam} core {cl}/{cm} total {tm}"
            if real != ours:
                bad += 1
                print(f"[MISMATCH oracle] {tag}: bowtie {len(real)} rows, oracle {len(ours)}; "
                      f"only bowtie {sorted(real - ours)[:5]}; only oracle {sorted(ours - real)[:5]}")
            else:
                print(f"[ok oracle]       {tag}: {len(real)} raw rows identical")
            if gpu is not None:
                from tests import util
                fixed = {21 + i for i, c in enumerate(pam) if c in "ACGT"}
                R = 20 + len(pam)
                valid, keys = set(), set()
                for name, strand, contig, off, desc in real:
                    keys.add(int(name))
                    mm = [int(d.split(":")[0]) for d in desc.split(",") if d]
                    if any((R - o) in fixed for o in mm):
                        continue
                    valid.add((int(name), contig, off, "+" if strand == "-" else "-"))
                hits, flags = gpu.offtarget_search(protos, pam=pam, core_length=cl, core_mismatches=cm, total_mismatches=tm)
                got = util.gpu_hit_set(gpu, hits)
                gkeys = set(np.nonzero(flags)[0].tolist())
                if got != valid or gkeys != keys:
                    bad += 1
                    print(f"[MISMATCH gpu]    {tag}: hits {len(got)} vs {len(valid)}, keys {len(gkeys)} vs {len(keys)}")
                else:
                    print(f"[ok gpu]          {tag}: {len(valid)} valid hits, {len(keys)} keys identical")
            golden["cases"].append({"pam": pam, "core_length": cl, "core_mismatches": cm, "total_mismatches": tm,
                                    "argv": argv[:-4] + ["-x", "IDX", "-f", "READS"], "tsv": tsv})
    if args.write_golden and not bad:
        out = os.path.join(ROOT, "tests", "golden", "bowtie_rows", f"bowtie131_seed{args.seed}.json.gz")
        with gzip.open(out, "wt") as fh:
            json.dump(golden, fh)
        print(f"wrote {out}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
