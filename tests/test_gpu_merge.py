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
        m.cap = 8  # far too small: the first pass only counts, the second is right-sized
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
        few = m.search(protos[:3])
        w3, _ = img.offtarget_search(protos[:3], want_raw_flags=False)
        assert np.array_equal(few, w3)
    finally:
        if created:
            dist.destroy_process_group()
