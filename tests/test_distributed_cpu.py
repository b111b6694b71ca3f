"""World-size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: the partition rule and the
merge of per-rank result lists.  The per-rank "search" is the oracle restricted to the rank's
shard, so the test checks that sharding + all-gather + canonical sort reproduces the 1-rank result."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from guido_b200 import _native
from tests import fakes, util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition():
    for padded in (64, 65536, 65537, 10_000_000, 3_200_000_000):
        for world in (1, 2, 3, 4, 8):
            prev, end = 0, (padded + 63) // 64 * 64
            for r in range(world):
                lo, hi = _native.shard_bounds(padded, r, world)
                assert lo == prev and (lo % 65536 == 0 or lo == end) and hi >= lo
                prev = hi
            assert prev == end
    with pytest.raises(ValueError):
        _native.shard_bounds(1000, 2, 2)


def _worker(rank, world, port, contigs, protos, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from guido_b200 import parallel
    dist.init_process_group("gloo", rank=rank, world_size=world)
    img = fakes.FakeImage(contigs)
    padded = img.contigs[-1][2] + ((img.contigs[-1][1] + 63) // 64) * 64 + 64
    lo, hi = parallel.shard_of(padded, rank, world)
    hits, flags = img.offtarget_search(protos)
    # a shard owns the hits whose PAM start lies in [lo, hi): '+' windows start 20 before their PAM
    pam_start = hits["gpos"].astype(np.int64) + np.where((hits["hi"] >> 31) == 0, 20, 0)
    mine = hits[(pam_start >= lo) & (pam_start < hi)]
    merged = parallel.all_gather_hits(mine)
    fw = np.array([p for p in range(lo, hi) if p % 7 == 0], dtype=np.uint32)  # any ascending per-shard list
    allpos = parallel.all_gather_positions(fw)
    if rank == 0:
        q.put((merged.tobytes(), allpos.tobytes(), padded))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_merge_equals_single_rank(world):
    contigs = util.synthetic_contigs(41, [90000, 70000], n_runs=2)
    s0 = contigs[0][1].upper()
    contigs[1] = (contigs[1][0], contigs[1][1][:1000] + s0[500:900] + contigs[1][1][1400:])
    rng = np.random.default_rng(9)
    protos = []
    while len(protos) < 12:
        p = int(rng.integers(500, 880))
        w = s0[p:p + 20]
        if not set(w) - set("ACGT"):
            protos.append(w)
    whole, _ = fakes.FakeImage(contigs).offtarget_search(protos)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, contigs, protos, q)) for r in range(world)]
    for p in procs:
        p.start()
    merged_bytes, pos_bytes, padded = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    merged = np.frombuffer(merged_bytes, dtype=_native.OT_HIT_DTYPE)
    assert len(whole) > 0 and np.array_equal(merged, whole)
    allpos = np.frombuffer(pos_bytes, dtype=np.uint32)
    total = (padded + 63) // 64 * 64
    assert np.array_equal(allpos, np.arange(0, total, 7, dtype=np.uint32))
