"""GPU parity of the raw kernels against the CPU oracle (through the C-ABI, via ctypes)."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

PAMS = ["NGG", "NAG", "TTTV", "NNGRRT", "YTN", "N", "GG", "NNNNGATT"]


@pytest.mark.parametrize("pam", PAMS)
def test_pam_scan_seq_matches_regex_oracle(native, oracle, pam):
    contigs = util.synthetic_contigs(11, [5000], n_runs=4, iupac=6)
    seq = contigs[0][1]
    fw, rv = native.pam_scan_seq(seq, pam)
    ofw, orv = oracle.pam_scan(seq, pam)
    assert np.array_equal(fw, ofw) and np.array_equal(rv, orv)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 22, 23, 63, 64, 65, 127, 128, 129, 16383, 16384, 16385, 40000])
def test_pam_scan_seq_sizes(native, oracle, n):
    rng = np.random.default_rng(n)
    seq = util.random_dna(rng, n)
    for pam in ("NGG", "GG", "TTTV"):
        fw, rv = native.pam_scan_seq(seq, pam)
        ofw, orv = oracle.pam_scan(seq, pam)
        assert np.array_equal(fw, ofw) and np.array_equal(rv, orv), (n, pam)


def test_pam_scan_bad_letter(native):
    with pytest.raises(KeyError):
        native.pam_scan_seq("ACGTACGT", "NGX")


def test_pam_scan_dense_tiles(native, oracle):
    # all-G sequence: every position matches NGG -> exercises the unstaged (dense) write path
    seq = "G" * 50000
    fw, rv = native.pam_scan_seq(seq, "NGG")
    ofw, orv = oracle.pam_scan(seq, "NGG")
    assert np.array_equal(fw, ofw) and np.array_equal(rv, orv)
    seq = "C" * 33000 + "G" * 33000
    fw, rv = native.pam_scan_seq(seq, "NGG")
    ofw, orv = oracle.pam_scan(seq, "NGG")
    assert np.array_equal(fw, ofw) and np.array_equal(rv, orv)


def _expected_genome_scan(oracle, contigs, offsets, pam, mode):
    """oracle positions -> global coordinates, with the mode's window filter applied on host"""
    P = len(pam)
    efw, erv = [], []
    for (name, seq), goff in zip(contigs, offsets):
        S = seq.upper()
        fw, rv = oracle.pam_scan(S, pam)
        for p in fw.tolist():
            if mode:
                w = S[p - 20:p + P] if p >= 20 else ""
                if len(w) != 20 + P:
                    continue
                if mode == 1 and "N" in w:
                    continue
                if mode == 2 and set(w) - set("ACGT"):
                    continue
            efw.append(goff + p)
        for p in rv.tolist():
            if mode:
                w = S[p:p + P + 20]
                if len(w) != 20 + P or set(w) - set("ACGT"):
                    continue
            erv.append(goff + p)
    return np.array(efw, np.uint32), np.array(erv, np.uint32)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("pam", ["NGG", "TTTV", "NAA"])
def test_pam_scan_genome_modes(native, oracle, pam, mode):
    contigs = util.synthetic_contigs(5, [30000, 777, 64, 20001, 10], n_runs=5, iupac=8)
    img = native.Image.from_seqs(contigs).upload()
    fw, rv = img.pam_scan_genome(pam, mode)
    efw, erv = _expected_genome_scan(oracle, contigs, [c[2] for c in img.contigs], pam, mode)
    assert np.array_equal(fw, efw)
    assert np.array_equal(rv, erv)


def test_pam_scan_region_equals_seq(native, oracle):
    contigs = util.synthetic_contigs(6, [50000, 3000], n_runs=4)
    img = native.Image.from_seqs(contigs).upload()
    for ci, a, b in [(0, 0, 50000), (0, 1234, 40001), (0, 49990, 50000), (1, 5, 5), (1, 100, 164), (0, 16384, 32768)]:
        fw, rv = img.pam_scan_region(ci, a, b, "NGG")
        ofw, orv = oracle.pam_scan(contigs[ci][1][a:b], "NGG")
        assert np.array_equal(fw, ofw) and np.array_equal(rv, orv), (ci, a, b)


def test_pam_scan_sharded_union(native):
    contigs = util.synthetic_contigs(8, [70000, 30000], n_runs=3)
    whole = native.Image.from_seqs(contigs).upload()
    fw, rv = whole.pam_scan_genome("NGG", 1)
    for n_shards in (2, 3, 8):
        parts = [native.Image.from_seqs(contigs).upload(shard_ix=i, n_shards=n_shards).pam_scan_genome("NGG", 1)
                 for i in range(n_shards)]
        assert np.array_equal(np.concatenate([p[0] for p in parts]), fw)
        assert np.array_equal(np.concatenate([p[1] for p in parts]), rv)


def _protos_from_genome(rng, contigs, n, mutate=3):
    """guides sampled from NGG sites of the genome, with a few random substitutions"""
    out = []
    seqs = [s.upper() for _, s in contigs]
    while len(out) < n:
        s = seqs[int(rng.integers(0, len(seqs)))]
        if len(s) < 40:
            continue
        p = int(rng.integers(0, len(s) - 23))
        w = s[p:p + 20]
        if set(w) - set("ACGT"):
            continue
        w = list(w)
        for _ in range(int(rng.integers(0, mutate + 1))):
            w[int(rng.integers(0, 20))] = "ACGT"[int(rng.integers(0, 4))]
        out.append("".join(w))
    return out


@pytest.mark.parametrize("pam,kw", [
    ("NGG", {}),
    ("NGG", dict(core_length=10, core_mismatches=2, total_mismatches=3)),
    ("NGG", dict(core_length=12, core_mismatches=1, total_mismatches=5)),
    ("NGG", dict(core_length=10, core_mismatches=0, total_mismatches=0)),
    ("NAG", {}),
    ("NRG", dict(core_mismatches=1)),
    ("TGG", dict(core_mismatches=3, total_mismatches=4)),
])
@pytest.mark.parametrize("algo", ["scan", "index", "table", "table-blocks"])
def test_offtarget_scan_matches_oracle(native, oracle, monkeypatch, pam, kw, algo):
    if algo == "table-blocks":  # bit-sliced compare of the join (only taken for core length 10)
        monkeypatch.setenv("GUIDO_OT_BLOCKS", "1")
    contigs = util.synthetic_contigs(21, [60000, 25000, 300], n_runs=4, iupac=5)
    rng = np.random.default_rng(3)
    protos = _protos_from_genome(rng, contigs, 60)
    img = native.Image.from_seqs(contigs).upload()
    hits, flags = img.offtarget_search(protos, pam=pam, algo={"scan": native.OT_SCAN, "index": native.OT_INDEX, "table": native.OT_TABLE,
                                             "table-blocks": native.OT_TABLE}[algo], **kw)
    valid, keys = util.oracle_valid_hits(oracle, contigs, protos, pam, **kw)
    assert util.gpu_hit_set(img, hits) == valid
    assert len(hits) == len(valid)
    assert set(np.nonzero(flags)[0].tolist()) == keys


def test_offtarget_hit_payload(native, oracle):
    """hi/lo planes of a hit decode to the genomic 23-mer in guide orientation"""
    contigs = util.synthetic_contigs(22, [40000], n_runs=2)
    rng = np.random.default_rng(4)
    protos = _protos_from_genome(rng, contigs, 40, mutate=2)
    img = native.Image.from_seqs(contigs).upload()
    hits, _ = img.offtarget_search(protos, algo=native.OT_SCAN)
    assert len(hits) > 0
    S = contigs[0][1].upper()
    ix, off = img.locate(hits["gpos"])
    for h, o in zip(hits, off):
        w = S[int(o):int(o) + 23]
        if int(h["hi"]) >> 31:
            w = oracle.rev_comp(w)
        dec = "".join("ACGT"[((int(h["lo"]) >> i) & 1) | (((int(h["hi"]) >> i) & 1) << 1)] for i in range(23))
        assert dec == w


def test_offtarget_sharded_union(native):
    contigs = util.synthetic_contigs(23, [50000, 50000], n_runs=3)
    rng = np.random.default_rng(5)
    protos = _protos_from_genome(rng, contigs, 50)
    whole, _ = native.Image.from_seqs(contigs).upload().offtarget_search(protos, algo=native.OT_SCAN)
    parts = []
    for i in range(4):
        h, _ = native.Image.from_seqs(contigs).upload(shard_ix=i, n_shards=4).offtarget_search(
            protos, algo=native.OT_SCAN, want_raw_flags=False)
        parts.append(h)
    merged = np.concatenate(parts)
    merged = merged[np.lexsort((merged["hi"] >> 31, merged["gpos"], merged["guide"]))]
    assert np.array_equal(merged, whole)


@pytest.mark.parametrize("core_length,core_mismatches", [(6, 2), (8, 0), (11, 1), (12, 2), (10, 2)])
def test_offtarget_index_equals_scan(native, core_length, core_mismatches):
    """the seed-index kernel returns byte-identical sorted hit lists to the brute-force kernel"""
    contigs = util.synthetic_contigs(24, [120000, 40000], n_runs=3)
    rng = np.random.default_rng(6)
    protos = _protos_from_genome(rng, contigs, 300, mutate=4) + ["ACGT" * 5, "G" * 20]
    protos += protos[:5]  # duplicated guides must each get their own hits
    img = native.Image.from_seqs(contigs).upload()
    kw = dict(core_length=core_length, core_mismatches=core_mismatches, total_mismatches=4, want_raw_flags=False)
    a, _ = img.offtarget_search(protos, algo=native.OT_SCAN, **kw)
    b, _ = img.offtarget_search(protos, algo=native.OT_INDEX, **kw)
    assert len(a) > 0 and np.array_equal(a, b)
    c, _ = img.offtarget_search(protos, algo=native.OT_TABLE, **kw)   # join with the PAM-site table
    assert np.array_equal(a, c)


def test_offtarget_index_rejects_unsupported_core(native):
    img = native.Image.from_seqs(util.synthetic_contigs(25, [5000], n_runs=0)).upload()
    with pytest.raises(ValueError, match="seed index"):
        img.offtarget_search(["ACGT" * 5], core_length=20, algo=native.OT_INDEX)
    hits, _ = img.offtarget_search(["ACGT" * 5], core_length=20, algo=native.OT_AUTO)  # falls back to the scan kernel


@pytest.mark.parametrize("wide", ["0", "1", "2"])
def test_offtarget_index_lookup_variants(native, monkeypatch, wide):
    """both bucket-walking strategies of the seed-index kernel give the brute-force answer"""
    monkeypatch.setenv("GUIDO_OT_WIDE", wide)
    contigs = util.synthetic_contigs(26, [150000], n_runs=3)
    rng = np.random.default_rng(7)
    protos = _protos_from_genome(rng, contigs, 500, mutate=4)
    img = native.Image.from_seqs(contigs).upload()
    a, _ = img.offtarget_search(protos, algo=native.OT_SCAN, want_raw_flags=False)
    b, _ = img.offtarget_search(protos, algo=native.OT_INDEX, want_raw_flags=False)
    assert len(a) > 0 and np.array_equal(a, b)


def test_site_table_lifecycle(native):
    """the PAM-site table is built once per (PAM sets, core length), reused, rebuilt on change,
    and a sharded image tables only its own shard"""
    contigs = util.synthetic_contigs(27, [200000, 70000], n_runs=4, iupac=3)
    rng = np.random.default_rng(8)
    protos = _protos_from_genome(rng, contigs, 200, mutate=4)
    img = native.Image.from_seqs(contigs).upload()
    fw, rv = img.pam_scan_genome("NGG", native.SCAN_STRICT)
    n, ms = img.site_table_build("NGG", 10)
    assert n == len(fw) + len(rv) and ms > 0
    assert img.site_table_build("NGG", 10) == (n, 0.0)            # reused
    ref, _ = img.offtarget_search(protos, algo=native.OT_SCAN, want_raw_flags=False)
    got, _ = img.offtarget_search(protos, algo=native.OT_TABLE, want_raw_flags=False)
    assert np.array_equal(ref, got)
    n8, ms8 = img.site_table_build("NGG", 8)                      # other core length: rebuilt
    assert n8 == n and ms8 > 0
    got, _ = img.offtarget_search(protos, pam="NAG", algo=native.OT_TABLE, want_raw_flags=False)
    ref, _ = img.offtarget_search(protos, pam="NAG", algo=native.OT_SCAN, want_raw_flags=False)
    assert np.array_equal(ref, got)                               # other PAM: rebuilt on the fly
    img.site_table_free()
    got, _ = img.offtarget_search(protos, pam="NAG", algo=native.OT_AUTO, want_raw_flags=False)
    assert np.array_equal(ref, got)
    parts = []
    for i in range(3):
        sh = native.Image.from_seqs(contigs).upload(shard_ix=i, n_shards=3)
        parts.append(sh.offtarget_search(protos, algo=native.OT_TABLE, want_raw_flags=False)[0])
    merged = np.concatenate(parts)
    merged = merged[np.lexsort((merged["hi"] >> 31, merged["gpos"], merged["guide"]))]
    whole, _ = img.offtarget_search(protos, algo=native.OT_SCAN, want_raw_flags=False)
    assert np.array_equal(merged, whole)


def test_site_table_empty_and_tiny(native):
    img = native.Image.from_seqs([("c", "ACGT" * 4)]).upload()      # shorter than a window
    assert img.site_table_build("NGG", 10)[0] == 0
    hits, flags = img.offtarget_search(["ACGT" * 5], algo=native.OT_TABLE)
    assert len(hits) == 0 and flags.tolist() == [0]
    img = native.Image.from_seqs([("c", "T" * 30 + "ACGTACGTACGTACGTACGTAGG" + "T" * 30)]).upload()
    hits, flags = img.offtarget_search(["ACGTACGTACGTACGTACGT", "GGGGGGGGGGCCCCCCCCCC"], algo=native.OT_TABLE)
    assert hits["guide"].tolist() == [0] and flags.tolist()[0] == 1


def test_offtarget_guides_are_packed_on_the_device(native):
    """letters travel as they are; packing and validation happen in a kernel"""
    img = native.Image.from_seqs([("c", "T" * 30 + "ACGTACGTACGTACGTACGTAGG" + "T" * 30)]).upload()
    with pytest.raises(ValueError, match=r"guide 1 holds a letter outside A/C/G/T at position 5"):
        img.offtarget_search(["ACGTACGTACGTACGTACGT", "ACGTNCGTACGTACGTACGT"])
    lower, _ = img.offtarget_search(["acgtacgtacgtacgtacgt"])
    upper, _ = img.offtarget_search(["ACGTACGTACGTACGTACGT"])
    assert len(upper) == 1 and np.array_equal(lower, upper)
    enc = native.encode_guides(["ACGTACGTACGTACGTACGT"])
    assert enc.tolist() == native.encode_guides(["acgtacgtacgtacgtacgt"]).tolist()


@pytest.mark.parametrize("pam", ["NGG", "TTTV", "TGGAC"])
def test_offtarget_packed_hits_expand(native, pam):
    """the 8-byte wire format + host-side expansion reproduces the 16-byte records bit for bit
    (both strands, windows across 64-base block borders, PAM lengths 3 / 4 / 5)"""
    contigs = util.synthetic_contigs(41, [150000, 40000, 513], n_runs=4, iupac=3)
    rng = np.random.default_rng(19)
    img = native.Image.from_seqs(contigs).upload()
    P = len(pam)
    # guides that do have sites for this PAM, on both strands: the protospacers next to real PAM hits
    fw, rv = img.pam_scan_genome(pam, native.SCAN_STRICT)
    comp = np.zeros(256, np.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    wf = img.fetch_windows(rng.choice(fw, 60, replace=False).astype(np.int64) - 20, 20)
    wr = comp[img.fetch_windows(rng.choice(rv, 60, replace=False).astype(np.int64) + P, 20)[:, ::-1]]
    protos = np.ascontiguousarray(np.concatenate([wf, wr]))
    full, f1 = img.offtarget_search(protos, pam=pam)
    packed, f2 = img.offtarget_search(protos, pam=pam, packed=True)
    assert packed.dtype == np.uint64 and len(packed) == len(full) >= 120 and np.array_equal(f1, f2)
    assert np.array_equal(packed >> np.uint64(33), full["guide"].astype(np.uint64))
    assert np.all(np.diff(packed.astype(np.uint64)) > 0)      # the packed word is the sort key
    assert (full["hi"] >> 31).any() and not (full["hi"] >> 31).all()
    assert np.array_equal(img.expand_hits(packed, P), full)


@pytest.mark.parametrize("total", [0, 2, 4, 6, 8])
@pytest.mark.parametrize("sliced", [True, False])
def test_offtarget_join_variants(native, monkeypatch, total, sliced):
    """the table join has a bit-sliced compare (core 10, total <= 7) and a popc compare: both must
    give the brute-force hit list, whatever the mismatch budget"""
    monkeypatch.setenv("GUIDO_OT_BLOCKS", "1" if sliced else "0")   # small guide sets default to popc
    contigs = util.synthetic_contigs(28, [180000, 20000], n_runs=3)
    rng = np.random.default_rng(9)
    protos = _protos_from_genome(rng, contigs, 400, mutate=5)
    img = native.Image.from_seqs(contigs).upload()
    kw = dict(core_length=10, core_mismatches=min(2, total), total_mismatches=total, want_raw_flags=False)
    a, _ = img.offtarget_search(protos, algo=native.OT_SCAN, **kw)
    b, _ = img.offtarget_search(protos, algo=native.OT_TABLE, **kw)
    assert len(a) > 0 and np.array_equal(a, b)


def test_site_table_capacity_fallback(native, monkeypatch):
    """when the PAM-site table cannot be held, AUTO streams the genome through the seed index
    (same hits); an explicit TABLE request says why it cannot run"""
    contigs = util.synthetic_contigs(29, [300000], n_runs=2)
    rng = np.random.default_rng(10)
    protos = _protos_from_genome(rng, contigs, 200, mutate=4)
    img = native.Image.from_seqs(contigs).upload()
    ref, _ = img.offtarget_search(protos, algo=native.OT_SCAN, want_raw_flags=False)
    monkeypatch.setenv("GUIDO_SITE_TABLE_MAX_MB", "0")
    got, _ = img.offtarget_search(protos, algo=native.OT_AUTO, want_raw_flags=False)
    assert len(ref) > 0 and np.array_equal(ref, got)
    with pytest.raises(MemoryError, match="PAM-site table does not fit"):
        img.offtarget_search(protos, algo=native.OT_TABLE, want_raw_flags=False)
    with pytest.raises(MemoryError, match="does not fit"):
        img.site_table_build("NGG", 10)
    monkeypatch.delenv("GUIDO_SITE_TABLE_MAX_MB")
    got, _ = img.offtarget_search(protos, algo=native.OT_TABLE, want_raw_flags=False)
    assert np.array_equal(ref, got)


@pytest.mark.parametrize("sliced", ["0", "1"])
def test_site_table_partition_union(native, monkeypatch, sliced):
    """key-range partition of the site table (the multi-GPU mode without replicated index builds):
    the parts' hit lists are disjoint and their union is the whole image's hit list"""
    monkeypatch.setenv("GUIDO_OT_BLOCKS", sliced)
    contigs = util.synthetic_contigs(30, [250000, 60000], n_runs=3)
    rng = np.random.default_rng(11)
    protos = _protos_from_genome(rng, contigs, 300, mutate=4)
    img = native.Image.from_seqs(contigs).upload()
    whole, _ = img.offtarget_search(protos, algo=native.OT_SCAN, want_raw_flags=False)
    n_all, _ = img.site_table_build("NGG", 10)
    for n_parts in (2, 8):
        parts, rows = [], 0
        for p in range(n_parts):
            img.site_table_partition(p, n_parts)
            rows += img.site_table_build("NGG", 10)[0]
            parts.append(img.offtarget_search(protos, algo=native.OT_AUTO, want_raw_flags=False)[0])
        assert rows == n_all
        merged = np.concatenate(parts)
        merged = merged[np.lexsort((merged["hi"] >> 31, merged["gpos"], merged["guide"]))]
        assert len(whole) > 0 and np.array_equal(merged, whole)
    with pytest.raises(ValueError, match="only the table join"):
        img.offtarget_search(protos, algo=native.OT_SCAN, want_raw_flags=False)
    with pytest.raises(ValueError, match="power of two"):
        img.site_table_partition(0, 3)
    img.site_table_partition(0, 1)
    back, _ = img.offtarget_search(protos, algo=native.OT_TABLE, want_raw_flags=False)
    assert np.array_equal(back, whole)


def test_offtarget_raw_flags_entry_point(native, oracle):
    """guido_offtarget_raw_flags = the key-presence half of the search on its own"""
    contigs = util.synthetic_contigs(31, [80000, 30000], n_runs=3, iupac=4)
    rng = np.random.default_rng(12)
    protos = _protos_from_genome(rng, contigs, 80, mutate=3)
    img = native.Image.from_seqs(contigs).upload()
    for pam, kw in (("NGG", {}), ("NAG", dict(core_mismatches=1, total_mismatches=3))):
        _, keys = util.oracle_valid_hits(oracle, contigs, protos, pam, **kw)
        flags = img.offtarget_raw_flags(protos, pam=pam, **kw)
        assert set(np.nonzero(flags)[0].tolist()) == keys
        # a key-partitioned handle answers the same: the pass ignores the partition
        img.site_table_partition(1, 4)
        assert np.array_equal(img.offtarget_raw_flags(protos, pam=pam, **kw), flags)
        img.site_table_partition(0, 1)
    assert len(img.offtarget_raw_flags([])) == 0


@pytest.mark.parametrize("sliced", ["0", "1"])
def test_offtarget_join_big_buckets(native, monkeypatch, sliced):
    """many copies of the same guides: buckets of several hundred keys, i.e. more than the two
    32-key blocks the sliced join evaluates per round (pointer-advance path), and long popc walks"""
    monkeypatch.setenv("GUIDO_OT_BLOCKS", sliced)
    contigs = util.synthetic_contigs(32, [200000], n_runs=2)
    rng = np.random.default_rng(13)
    img = native.Image.from_seqs(contigs).upload()
    fw, _ = img.pam_scan_genome("NGG", native.SCAN_STRICT)     # guides that do have sites: real NGG windows
    S, off = contigs[0][1].upper(), img.contigs[0][2]
    base = [S[int(p) - off - 20:int(p) - off] for p in rng.choice(fw, size=40, replace=False)]
    protos = base[:2] * 170 + base[2:] + base[:1] * 37      # 340 + 38 + 37 guides, heavy duplication
    a, fa = img.offtarget_search(protos, algo=native.OT_SCAN)
    b, fb = img.offtarget_search(protos, algo=native.OT_TABLE)
    assert len(a) > 300 and np.array_equal(a, b) and np.array_equal(fa, fb)
    img.site_table_partition(1, 2)
    c, _ = img.offtarget_search(protos, algo=native.OT_TABLE, want_raw_flags=False)
    img.site_table_partition(0, 2)
    d, _ = img.offtarget_search(protos, algo=native.OT_TABLE, want_raw_flags=False)
    merged = np.concatenate([c, d])
    merged = merged[np.lexsort((merged["hi"] >> 31, merged["gpos"], merged["guide"]))]
    assert np.array_equal(merged, a)


@pytest.mark.parametrize("cm,total", [(2, 4), (1, 3), (2, 2), (0, 2)])
def test_index_build_gather_equals_scatter(native, monkeypatch, cm, total):
    """the block join's seed index is built bucket-major by gathering the guides of the occupied
    neighbour cores (default) or by the atomic scatter (GUIDO_IDX_BUILD=scatter, and core budgets the
    gather does not cover): same hit list, also with cores shared by many guides and with a
    partitioned key space"""
    monkeypatch.setenv("GUIDO_OT_BLOCKS", "1")
    contigs = util.synthetic_contigs(41, [260000, 30000], n_runs=3)
    rng = np.random.default_rng(23)
    img = native.Image.from_seqs(contigs).upload()
    fw, _ = img.pam_scan_genome("NGG", native.SCAN_STRICT)     # guides that do have sites: real NGG windows, mutated
    S, off = contigs[0][1].upper(), img.contigs[0][2]
    protos = []
    for p in rng.choice(fw[fw < off + len(S)], size=700, replace=False):
        w = list(S[int(p) - off - 20:int(p) - off])
        for _ in range(int(rng.integers(0, 4))):
            w[int(rng.integers(0, 20))] = "ACGT"[int(rng.integers(0, 4))]
        protos.append("".join(w))
    protos = protos + protos[:3] * 40 + ["ACGTACGT" + p[8:] for p in protos[:50]]   # shared cores, other distal bases
    kw = dict(core_length=10, core_mismatches=cm, total_mismatches=total, want_raw_flags=False)
    ref, _ = img.offtarget_search(protos, algo=native.OT_SCAN, **kw)
    out = {}
    for mode in ("gather", "scatter"):
        monkeypatch.setenv("GUIDO_IDX_BUILD", mode)
        out[mode], _ = img.offtarget_search(protos, algo=native.OT_TABLE, **kw)
        parts = []
        for ix in range(4):
            img.site_table_partition(ix, 4)
            parts.append(img.offtarget_search(protos, algo=native.OT_TABLE, **kw)[0])
        img.site_table_partition(0, 1)
        merged = np.concatenate(parts)
        merged = merged[np.lexsort((merged["hi"] >> 31, merged["gpos"], merged["guide"]))]
        assert np.array_equal(merged, out[mode])
    assert len(ref) > 100
    assert np.array_equal(out["scatter"], ref)
    assert np.array_equal(out["gather"], ref)


@pytest.mark.parametrize("total", [3, 4, 5])
def test_offtarget_block_join_full_passes(native, oracle, monkeypatch, total):
    """dense buckets on both sides (2 000 guides x 40 copies, a 24 Mb image): the block join runs
    whole 32-entry passes -- the short "at most 2 of 10" tree when every lane carries that bias
    (total 4), the general tree otherwise -- plus shared remainder passes; the result must equal
    the popc join's, and the oracle's on a sample of the guides"""
    img = native.Image.synth([15_000_000, 9_000_000], seed=91, n_frac=0.004).upload()
    rng = np.random.default_rng(17)
    fw, _ = img.pam_scan_genome("NGG", native.SCAN_STRICT)
    pos = rng.choice(fw, size=2000, replace=False).astype(np.int64)
    base = img.fetch_windows(pos - 20, 20)
    mut = rng.integers(0, 20, size=len(base))
    base[np.arange(len(base)), mut] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=len(base))]
    protos = np.ascontiguousarray(np.tile(base, (40, 1)))
    kw = dict(core_length=10, core_mismatches=2, total_mismatches=total, want_raw_flags=False)
    monkeypatch.setenv("GUIDO_OT_BLOCKS", "0")
    a, _ = img.offtarget_search(protos, algo=native.OT_TABLE, **kw)
    monkeypatch.setenv("GUIDO_OT_BLOCKS", "1")
    b, _ = img.offtarget_search(protos, algo=native.OT_TABLE, **kw)
    assert len(a) > 40 * 2000 and np.array_equal(a, b)
    # every copy of a guide owns the same hits
    first = b[b["guide"] < 2000]
    for c in (1, 17, 39):
        cp = b[(b["guide"] >= 2000 * c) & (b["guide"] < 2000 * (c + 1))]
        assert np.array_equal(cp["gpos"], first["gpos"]) and np.array_equal(cp["hi"], first["hi"])
    # oracle on a few guides against the part of the genome that holds their hits is too slow at
    # 24 Mb; brute force on the device is the independent check here
    sub = np.ascontiguousarray(base[:64])
    c, _ = img.offtarget_search(sub, algo=native.OT_SCAN, **kw)
    assert np.array_equal(c, first[first["guide"] < 64])


@pytest.mark.parametrize("pam,mode", [("NGG", 0), ("NGG", 2), ("TTTV", 1), ("N", 0), ("NNGRRT", 2)])
def test_pam_scan_one_pass_equals_two_pass(native, monkeypatch, pam, mode):
    """The single-pass (look-back) kernel and the count -> fill pair must return the same lists:
    ~700 tiles (several look-back rounds per tile), N runs across tile borders, and 'N' as the PAM
    (every position a hit: the dense, unstaged write path of both kernels)."""
    img = native.Image.synth([30_000_000, 12_345_678, 3_000_001, 70_000], seed=77, n_frac=0.01).upload()
    monkeypatch.setenv("GUIDO_PAM_PASSES", "2")
    fw2, rv2 = img.pam_scan_genome(pam, mode)
    monkeypatch.setenv("GUIDO_PAM_PASSES", "1")
    for _ in range(3):  # repeated: descriptors and the ticket are reset per launch
        fw1, rv1 = img.pam_scan_genome(pam, mode)
        assert np.array_equal(fw1, fw2) and np.array_equal(rv1, rv2)
    assert len(fw2) > 100000 and np.all(np.diff(fw2.astype(np.int64)) > 0)
    # capacity too small: exact counts come back, nothing is written past the caller's buffers
    import ctypes
    cap = len(fw2) // 2
    fw, rv = np.full(cap + 8, 0xdeadbeef, np.uint32), np.full(cap + 8, 0xdeadbeef, np.uint32)
    nf, nr = ctypes.c_int64(), ctypes.c_int64()
    rc = native.lib().guido_pam_scan_genome(img._h, pam.encode(), mode, fw.ctypes.data, cap, ctypes.byref(nf),
                                            rv.ctypes.data, cap, ctypes.byref(nr), None)
    assert rc == native.ERR_CAPACITY and nf.value == len(fw2) and nr.value == len(rv2)
    assert np.all(fw[cap:] == 0xdeadbeef) and np.all(rv[cap:] == 0xdeadbeef)


@pytest.mark.parametrize("n_guides", [0, 1, 333])
def test_offtarget_csr_format(native, monkeypatch, n_guides):
    """CSR by guide (positions + strand bits + per-guide offsets) holds exactly the packed hit list,
    including guides without hits at either end and an empty guide set"""
    contigs = util.synthetic_contigs(52, [220000, 40000], n_runs=3)
    rng = np.random.default_rng(29)
    img = native.Image.from_seqs(contigs).upload()
    protos = _protos_from_genome(rng, contigs, n_guides, mutate=2) if n_guides else []
    if n_guides > 10:
        protos[0] = "ACGTTGCAACGTTGCAACGT"      # (most sampled guides have no NGG site anyway: many empty segments)
        protos[-1] = "TTTTTTTTTTGGGGGGGGGG"
    for algo in (native.OT_TABLE, native.OT_SCAN):
        packed, f1 = img.offtarget_search(protos, algo=algo, packed=True)
        offs, gpos, bits, f2 = img.offtarget_search_csr(protos, algo=algo)
        assert offs[0] == 0 and offs[-1] == len(packed) == len(gpos) and np.all(np.diff(offs) >= 0)
        assert np.array_equal(native.Image.csr_to_packed(offs, gpos, bits), packed)
        assert np.array_equal(f1, f2)
    if n_guides > 10:
        assert len(packed) > 0
        # the 8-byte records are sorted by the segmented sort or, for guides with very long lists, by the
        # key-only radix sort (GUIDO_SORT=radix forces it): same list; and it is the full records' order
        monkeypatch.setenv("GUIDO_SORT", "radix")
        p2, _ = img.offtarget_search(protos, algo=native.OT_TABLE, packed=True)
        o2, g2, b2, _ = img.offtarget_search_csr(protos, algo=native.OT_TABLE)
        monkeypatch.delenv("GUIDO_SORT")
        full, _ = img.offtarget_search(protos, algo=native.OT_TABLE)
        assert np.array_equal(p2, packed) and np.array_equal(o2, offs) and np.array_equal(g2, gpos) and np.array_equal(b2, bits)
        assert np.array_equal(img.expand_hits(packed), full)
