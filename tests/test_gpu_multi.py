"""Multi-GPU path on real devices: torchrun-style NCCL run of ShardedGenome on 2 GPUs must return
the single-GPU hit list byte for byte.  Skipped on a one-GPU box (the gloo tests cover the logic)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from guido_b200 import _native
from tests import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from guido_b200 import _native as N, parallel
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
d = np.load(sys.argv[2], allow_pickle=True)
sg = parallel.ShardedGenome(N.Image.load(str(d["image"])), partition=sys.argv[4])
hits, flags = sg.offtarget_search(list(d["protos"]))
fw, rv = sg.pam_scan("NGG")
if rank == 0:
    np.savez(sys.argv[3], hits=hits.view(np.uint32), flags=flags, fw=fw, rv=rv)
dist.barrier(); dist.destroy_process_group()
'''


@pytest.mark.skipif(_native.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("partition", ["position", "key"])
def test_two_gpu_sharded_search_equals_one_gpu(tmp_path, partition):
    contigs = util.synthetic_contigs(51, [400000, 300000], n_runs=3)
    img = _native.Image.from_seqs(contigs)
    img.save(tmp_path / "g.g2b")
    rng = np.random.default_rng(11)
    s0 = contigs[0][1].upper()
    protos = []
    while len(protos) < 64:
        p = int(rng.integers(0, len(s0) - 23))
        if not set(s0[p:p + 20]) - set("ACGT"):
            protos.append(s0[p:p + 20])
    img.upload(device=0)
    whole, wflags = img.offtarget_search(protos)
    wfw, wrv = img.pam_scan_genome("NGG")
    np.savez(tmp_path / "in.npz", image=str(tmp_path / "g.g2b"), protos=np.array(protos))
    (tmp_path / "worker.py").write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533" if partition == "position" else "29534", str(tmp_path / "worker.py"), ROOT,
           str(tmp_path / "in.npz"), str(tmp_path / "out.npz"), partition]
    subprocess.run(cmd, check=True, timeout=300)
    out = np.load(tmp_path / "out.npz")
    assert np.array_equal(out["hits"].view(_native.OT_HIT_DTYPE).reshape(-1), whole)
    assert np.array_equal(out["flags"], wflags)
    assert np.array_equal(out["fw"], wfw) and np.array_equal(out["rv"], wrv)


PEER_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from guido_b200 import parallel
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"])); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n_slots, words = 3, 1000
pg = parallel.PeerGather(dev, n_slots, words)
ok = True
for rep in range(3):
    src = [torch.arange(words, dtype=torch.int64, device=dev) + 1000000 * rank + 1000 * k + rep for k in range(n_slots)]
    torch.cuda.synchronize()
    ready = torch.cuda.Event(); ready.record(torch.cuda.current_stream())
    for k in range(n_slots):
        pg.send(k, src[k].data_ptr(), after=ready if k % 2 else None)
    pg.fence()
    pg.stream.synchronize()
    got = torch.empty(pg.words, dtype=torch.int64, device=dev)
    from guido_b200 import _native as N
    N.check(N.lib().guido_copy_dev(got.data_ptr(), pg.local, pg.words * 8, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    got = got.cpu().numpy().reshape(n_slots, world, words)
    for k in range(n_slots):
        for r in range(world):
            ok &= bool(np.array_equal(got[k, r], np.arange(words) + 1000000 * r + 1000 * k + rep))
    dist.barrier()
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    open(sys.argv[2], "w").write("ok" if int(flag.item()) else "mismatch")
dist.barrier(); pg.close(); dist.destroy_process_group()
'''


@pytest.mark.skipif(_native.device_count() < 2, reason="needs 2 GPUs")
def test_peer_gather_fills_every_ranks_buffer(tmp_path):
    """slots copied into the other ranks' receive buffers (CUDA IPC) land where an all-gather would put them"""
    (tmp_path / "worker.py").write_text(PEER_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29535", str(tmp_path / "worker.py"), ROOT, str(tmp_path / "out.txt")]
    subprocess.run(cmd, check=True, timeout=300)
    assert (tmp_path / "out.txt").read_text() == "ok"
