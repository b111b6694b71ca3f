"""Device-resident search + merge (guido_offtarget_search_dev, guido_hits_sort_dev,
parallel.DeviceMerge): the hit list that never leaves the GPU before it is complete must equal the
host-buffer call's byte for byte."""
import numpy as np
import pytest

from tests import util
from tests.test_gpu_kernels import _protos_from_genome

pytestmark = pytest.mark.gpu


def _case(native, seed=23, n=300):
    contigs = util.synthetic_contigs(seed, [300000, 150000], n_runs=3)
    rng = np.random.default_rng(seed)
    protos = _protos_from_genome(rng, contigs, n)
    img = native.Image.from_seqs(contigs).upload()
    return img, protos


@pytest.mark.parametrize("algo", ["table", "index", "scan"])
def test_search_dev_then_sort_dev_equals_host_call(native, algo):
    import torch
    img, protos = _case(native)
    a = {"table": native.OT_TABLE, "index": native.OT_INDEX, "scan": native.OT_SCAN}[algo]
    want, _ = img.offtarget_search(protos, algo=a, want_raw_flags=False)
    dev = torch.device("cuda", int(img.info.device))
    enc = native.encode_guides(protos)
    d_guides = torch.from_numpy(enc.view(np.int32)).to(dev)
    cap = len(want) + 64
    d_hits = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    d_count = torch.zeros(1, dtype=torch.int64, device=dev)
    s = torch.cuda.current_stream(dev).cuda_stream
    img.offtarget_search_dev(d_guides.data_ptr(), len(protos), "NGG", 10, 2, 4, a, d_hits.data_ptr(), cap,
                             d_count.data_ptr(), s)
    n = int(d_count.item())
    assert n == len(want)
    img.sort_hits_dev(d_hits.data_ptr(), n, len(protos), s)
    got = d_hits[:n].cpu().numpy().view(native.OT_HIT_DTYPE).reshape(-1)
    assert np.array_equal(got, want)
    # sorting a sorted list, an empty list and a single record are no-ops
    img.sort_hits_dev(d_hits.data_ptr(), n, len(protos), s)
    img.sort_hits_dev(d_hits.data_ptr(), 0, len(protos), s)
    img.sort_hits_dev(d_hits.data_ptr(), 1, len(protos), s)
    assert np.array_equal(d_hits[:n].cpu().numpy().view(native.OT_HIT_DTYPE).reshape(-1), want)
    with pytest.raises(Exception):
        img.sort_hits_dev(d_hits.data_ptr(), -1, len(protos), s)


def test_device_merge_single_rank_group(native):
    """DeviceMerge on a one-rank NCCL group (the 2-GPU run is tests/test_gpu_multi.py): capacity
    growth, repeated calls on kept buffers, letters given as strings or as a uint8 array."""
    import torch
    import torch.distributed as dist
    from guido_b200.parallel import DeviceMerge
    img, protos = _case(native, seed=29, n=200)
    want, _ = img.offtarget_search(protos, want_raw_flags=False)
    torch.cuda.set_device(int(img.info.device))
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29571", rank=0, world_size=1,
                                device_id=torch.device("cuda", int(img.info.device)))
    try:
        m = DeviceMerge(img)
        m.stride = 9  # far too small: the first pass only counts, the second is right-sized
        got = m.search(protos).copy()
        assert np.array_equal(got, want)
        as_u8 = np.frombuffer("".join(protos).encode(), np.uint8).reshape(-1, 20)
        assert np.array_equal(m.search(as_u8), want)
        t = m.search(protos, to_host=False)
        assert t.is_cuda and tuple(t.shape) == (len(want), 4)
        with pytest.raises(ValueError, match="guide 1 holds a letter outside A/C/G/T at position 5"):
            m.search([protos[0], protos[1][:4] + "N" + protos[1][5:]])
        with pytest.raises(ValueError, match="20 nt"):
            m.search(["ACGT"])
        packed = m.search(protos, to_host="shared")  # 8-byte records in the node-shared pinned segment
        want_packed, _ = img.offtarget_search(protos, want_raw_flags=False, packed=True)
        assert packed.dtype == np.uint64 and np.array_equal(packed, want_packed)
        few = m.search(protos[:3])
        w3, _ = img.offtarget_search(protos[:3], want_raw_flags=False)
        assert np.array_equal(few, w3)
    finally:
        if created:
            dist.destroy_process_group()


def _random_hits(rng, n, n_guides, heavy=None):
    """n distinct (guide, gpos, strand) records in random order; `heavy` = (guide, count) gives one
    guide a long list"""
    g = rng.integers(0, n_guides, n).astype(np.uint32)
    if heavy is not None:
        g[:heavy[1]] = heavy[0]
    gpos = rng.permutation(np.arange(n, dtype=np.uint32) * 7 + 3)  # distinct positions
    strand = rng.integers(0, 2, n).astype(np.uint32)
    hits = np.zeros(n, dtype=[("guide", "<u4"), ("gpos", "<u4"), ("hi", "<u4"), ("lo", "<u4")])
    hits["guide"], hits["gpos"] = g, gpos
    hits["hi"] = (strand << 31) | rng.integers(0, 1 << 23, n).astype(np.uint32)
    hits["lo"] = rng.integers(0, 1 << 23, n).astype(np.uint32)
    return hits[rng.permutation(n)]


@pytest.mark.parametrize("n,n_guides,heavy", [
    (2, 1, None), (1000, 1, None), (5000, 7, None), (200000, 1000, None), (300000, 100000, None),
    (100000, 50, (3, 4096)),      # largest segment the rank sort takes (keys read through L1)
    (100000, 50, (3, 4097)),      # one guide above it: general radix path
    (50000, 1 << 20, None),       # far more guides than hits: mostly empty segments
])
def test_sort_hits_dev_matches_lexsort(native, monkeypatch, n, n_guides, heavy):
    import torch
    from guido_b200.parallel import canonical_order
    rng = np.random.default_rng(n + n_guides)
    hits = _random_hits(rng, n, n_guides, heavy)
    want = canonical_order(hits)
    img = native.Image.from_seqs([("c", "ACGT" * 64)]).upload()
    dev = torch.device("cuda", int(img.info.device))
    s = torch.cuda.current_stream(dev).cuda_stream
    for mode in (None, "radix"):
        if mode:
            monkeypatch.setenv("GUIDO_SORT", mode)
        d = torch.from_numpy(hits.view(np.int32).reshape(-1, 4).copy()).to(dev)
        img.sort_hits_dev(d.data_ptr(), n, n_guides, s)
        got = d.cpu().numpy().view(np.uint32).reshape(-1).view(want.dtype)
        assert np.array_equal(got, want), mode


@pytest.mark.parametrize("sliced", ["0", "1"])
def test_join_schedules_agree(native, monkeypatch, sliced):
    """The join launched slice by slice next to the index build (default) and as one launch after
    the build (GUIDO_OT_OVERLAP=0, bench.py's roofline pass) return the same list; so do several
    index slices (GUIDO_IDX_SLICE_MB=1, small enough to cut this index into the maximum of 32)."""
    monkeypatch.setenv("GUIDO_OT_BLOCKS", sliced)
    img, protos = _case(native, seed=31, n=400)
    want, wf = img.offtarget_search(protos, algo=native.OT_SCAN)
    for env in ({}, {"GUIDO_OT_OVERLAP": "0"}, {"GUIDO_IDX_SLICE_MB": "1"},
                {"GUIDO_IDX_SLICE_MB": "1", "GUIDO_OT_OVERLAP": "0"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for _ in range(2):  # back to back: events and streams are reused
            got, gf = img.offtarget_search(protos, algo=native.OT_TABLE)
            assert np.array_equal(got, want) and np.array_equal(gf, wf), env
        for k in env:
            monkeypatch.delenv(k)


@pytest.mark.parametrize("cl,cm,tm", [(10, 2, 4), (10, 1, 3), (10, 3, 4), (9, 2, 4), (7, 2, 3), (6, 1, 2), (10, 0, 2), (11, 2, 4)])
def test_bucket_counts_by_convolution_equal_histogram(native, monkeypatch, cl, cm, tm):
    """Bucket sizes of the guide index from the Hamming-ball convolution (default for cores up to 10
    bases) and from the histogram pass (one atomic per guide and variant) give the same index:
    identical hit lists for odd and even core lengths, every seed budget, key-range parts included.
    (pam NGG has one N: core_mismatches + 1 <= 3.)"""
    if cm + 1 > 3:
        pytest.skip("bowtie's -n is at most 3")
    img, protos = _case(native, seed=37 + cl, n=500)
    protos = protos + protos[:50]  # repeated guides: buckets with several entries of the same key word
    kw = dict(core_length=cl, core_mismatches=cm, total_mismatches=tm)
    want, wf = img.offtarget_search(protos, algo=native.OT_SCAN, **kw)
    for mode in (None, "atomic"):
        if mode:
            monkeypatch.setenv("GUIDO_IDX_HIST", mode)
        for algo in (native.OT_TABLE, native.OT_INDEX):
            got, gf = img.offtarget_search(protos, algo=algo, **kw)
            assert np.array_equal(got, want) and np.array_equal(gf, wf), (mode, algo)
    monkeypatch.delenv("GUIDO_IDX_HIST")
    if cl == 10 and cm == 2:
        n_all, _ = img.site_table_build("NGG", cl)
        parts = []
        for p in range(4):
            img.site_table_partition(p, 4)
            parts.append(img.offtarget_search(protos, algo=native.OT_TABLE, want_raw_flags=False, **kw)[0])
        img.site_table_partition(0, 1)
        merged = np.concatenate(parts)
        merged = merged[np.lexsort((merged["hi"] >> 31, merged["gpos"], merged["guide"]))]
        assert np.array_equal(merged, want)


def test_packed_slots_concat(native):
    """merge on the 8-byte wire format: pack behind a header word (count read on the device), slots
    laid out as an all-gather would leave them, concat -> the packed form of the host call's list;
    a slot that is too small is cut and shows in its header"""
    import torch
    img, protos = _case(native, seed=31, n=400)
    want, _ = img.offtarget_search(protos, algo=native.OT_TABLE, want_raw_flags=False, packed=True)
    dev = torch.device("cuda", int(img.info.device))
    s = torch.cuda.current_stream(dev).cuda_stream
    d_guides = torch.from_numpy(native.encode_guides(protos).view(np.int32)).to(dev)
    parts, counts = [], []
    for ix in range(2):           # the two halves of the key space, as two ranks would hold them
        img.site_table_partition(ix, 2)
        d_hits = torch.empty((len(want) + 8, 4), dtype=torch.int32, device=dev)
        d_count = torch.zeros(1, dtype=torch.int64, device=dev)
        img.offtarget_search_dev(d_guides.data_ptr(), len(protos), "NGG", 10, 2, 4, native.OT_TABLE, d_hits.data_ptr(),
                                 d_hits.shape[0], d_count.data_ptr(), s)
        parts.append((d_hits, d_count)); counts.append(int(d_count.item()))
    img.site_table_partition(0, 1)
    assert sum(counts) == len(want) and min(counts) > 0
    for stride in (max(counts) + 1, max(counts) + 50, min(counts) // 2):
        d_gath = torch.full((2 * stride,), -1, dtype=torch.int64, device=dev)
        for r, (d_hits, d_count) in enumerate(parts):
            img.pack_send_dev(d_hits.data_ptr(), d_count.data_ptr(), stride - 1, d_gath[r * stride:].data_ptr(), s)
        d_out = torch.empty(2 * stride, dtype=torch.int64, device=dev)
        d_total = torch.zeros(1, dtype=torch.int64, device=dev)
        img.concat_packed_dev(d_gath.data_ptr(), 2, stride, d_out.data_ptr(), d_out.shape[0], d_total.data_ptr(), s)
        heads = d_gath[::stride].cpu().numpy()
        assert heads.tolist() == counts
        total = int(d_total.item())
        if stride > max(counts):
            assert total == len(want)
            assert np.array_equal(np.sort(d_out[:total].cpu().numpy().view(np.uint64)), want)
        else:
            assert total == 2 * (stride - 1)   # both parts cut: the caller sees counts > stride - 1 in the headers


@pytest.mark.parametrize("n_slices", [1, 2, 4])
def test_sliced_search_union_and_events(native, monkeypatch, n_slices):
    """the key range searched slice by slice (hits, count and event per slice) = the search in one go;
    also on a partitioned handle and when the block join is not applicable (everything in slice 0)"""
    import torch
    img, protos = _case(native, seed=37, n=500)
    dev = torch.device("cuda", int(img.info.device))
    s = torch.cuda.current_stream(dev)
    d_guides = torch.from_numpy(native.encode_guides(protos).view(np.int32)).to(dev)
    for blocks, part in (("1", None), ("1", (1, 2)), ("0", None)):
        monkeypatch.setenv("GUIDO_OT_BLOCKS", blocks)
        if part:
            img.site_table_partition(*part)
        want, _ = img.offtarget_search(protos, algo=native.OT_TABLE, want_raw_flags=False)
        cap = len(want) + 16
        d_hits = torch.zeros((n_slices * cap, 4), dtype=torch.int32, device=dev)
        d_counts = torch.full((n_slices,), -1, dtype=torch.int64, device=dev)
        events = [torch.cuda.Event() for _ in range(n_slices)]
        for e in events:
            e.record(s)     # (creates the handle)
        img.offtarget_search_sliced_dev(d_guides.data_ptr(), len(protos), "NGG", 10, 2, 4, n_slices, d_hits.data_ptr(), cap,
                                        d_counts.data_ptr(), [e.cuda_event for e in events], s.cuda_stream)
        for e in events:
            e.synchronize()
        counts = d_counts.cpu().numpy()
        assert counts.sum() == len(want) and (counts >= 0).all()
        if blocks == "1" and n_slices > 1:
            assert (counts > 0).all()
        got = np.concatenate([d_hits[k * cap:k * cap + int(c)].cpu().numpy().view(native.OT_HIT_DTYPE).reshape(-1)
                              for k, c in enumerate(counts)])
        got = got[np.lexsort((got["hi"] >> 31, got["gpos"], got["guide"]))]
        assert np.array_equal(got, want)
        img.site_table_partition(0, 1)
