"""Shared helpers for the tests: seeded synthetic genomes and oracle adaptors."""
import numpy as np

LETTERS = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_dna(rng, n):
    return LETTERS[rng.integers(0, 4, size=n)].tobytes().decode()


def synthetic_contigs(seed, lens, n_runs=3, run_len=(5, 200), lower_frac=0.1, iupac=0):
    """[(name, seq)] uniform ACGT with N runs, some lower-case blocks, optional stray IUPAC codes."""
    rng = np.random.default_rng(seed)
    out = []
    for ci, n in enumerate(lens):
        s = bytearray(random_dna(rng, n).encode())
        for _ in range(n_runs):
            if n < 50:
                break
            ln = int(rng.integers(run_len[0], run_len[1] + 1))
            st = int(rng.integers(0, max(1, n - ln)))
            s[st:st + ln] = b"N" * min(ln, n - st)
        for _ in range(iupac):
            s[int(rng.integers(0, n))] = int(rng.choice(list(b"RYSWKMBDHV")))
        blk = 100
        for b in range(0, n, blk):
            if rng.random() < lower_frac:
                s[b:b + blk] = bytes(s[b:b + blk]).lower()
        out.append((f"chr{ci + 1}", s.decode()))
    return out


def oracle_valid_hits(O, contigs, protos, pam, core_length=10, core_mismatches=2, total_mismatches=4):
    """Oracle -> (set of (guide, contig_name, offset, guido_strand) for PAM-valid alignments,
    set of guides with any raw alignment)."""
    seed_mms, seed_len, e = O.bowtie_args(pam, core_length, core_mismatches, total_mismatches)
    reads = [O.rev_comp(p + pam) for p in protos]
    raw = O.bowtie_n(contigs, reads, seed_mms, seed_len, e)
    R = 20 + len(pam)
    fixed = {21 + i for i, c in enumerate(pam) if c in "ACGT"}
    valid, keys = set(), set()
    for h in raw:
        keys.add(h["read_ix"])
        if any((R - off) in fixed for off, _, _ in h["mm"]):
            continue
        valid.add((h["read_ix"], h["contig"], h["offset"], "+" if h["strand"] == "-" else "-"))
    return valid, keys


def gpu_hit_set(img, hits):
    ix, off = img.locate(hits["gpos"])
    names = [c[0] for c in img.contigs]
    return {(int(g), names[int(c)], int(o), "-" if (int(h) >> 31) else "+")
            for g, c, o, h in zip(hits["guide"], ix, off, hits["hi"])}
