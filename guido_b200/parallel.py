"""Multi-GPU data parallelism for the hot path: one process per GPU (torchrun), genome sharded.

The reference has no parallel code (its only concurrency is bowtie's ``-p``).  Here every window
of the genome is independent, so the padded genome coordinate is cut into ``world_size``
contiguous, 65536-base aligned chunks (``guido_shard_bounds``); rank r uploads chunk r plus a
256-base halo and owns the PAM starts inside it, so no window is reported twice and none is lost.
Guides are replicated.  The only exchange is the merge of the per-rank result lists:

* off-target hits: all-gather of the counts, then of the (padded) 16-byte records -- NCCL over
  NVLink when the tensors live on the GPUs, gloo for the CPU tests;
* PAM positions: the per-rank lists are globally ordered by construction (rank order = genome
  order), so the merge is a concatenation.

A final canonical sort by (guide, global position, strand) makes 1/2/4/8-GPU outputs byte-identical.
"""
import numpy as np

from . import _native


def shard_of(padded_bases, rank, world):
    return _native.shard_bounds(padded_bases, rank, world)


def canonical_order(hits):
    """sort a hit array by (guide, gpos, strand) -- the order guido_offtarget_search returns"""
    if len(hits) == 0:
        return hits
    return hits[np.lexsort((hits["hi"] >> 31, hits["gpos"], hits["guide"]))]


def all_gather_hits(hits, group=None, device=None):
    """Every rank contributes its hit array (OT_HIT_DTYPE); every rank gets the merged, canonically
    ordered list.  Works on whatever backend the default process group uses."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    local = torch.from_numpy(np.ascontiguousarray(hits).view(np.int32).reshape(-1, 4).copy()).to(dev)
    count = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(max(counts), 1)
    padded = torch.zeros((m, 4), dtype=torch.int32, device=dev)
    padded[:local.shape[0]] = local
    parts = [torch.zeros((m, 4), dtype=torch.int32, device=dev) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    merged = np.concatenate([p[:c].cpu().numpy() for p, c in zip(parts, counts)]) if sum(counts) else \
        np.zeros((0, 4), np.int32)
    return canonical_order(np.ascontiguousarray(merged).view(_native.OT_HIT_DTYPE).reshape(-1))


def all_gather_positions(positions, group=None, device=None):
    """Concatenate per-rank ascending position lists in rank order (= genome order)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    local = torch.from_numpy(np.ascontiguousarray(positions).astype(np.int64)).to(dev)
    count = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(max(counts), 1)
    padded = torch.zeros(m, dtype=torch.int64, device=dev)
    padded[:local.shape[0]] = local
    parts = [torch.zeros(m, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return np.concatenate([p[:c].cpu().numpy() for p, c in zip(parts, counts)]).astype(np.uint32)


class ShardedGenome:
    """This rank's shard of a genome image, resident on this rank's GPU."""

    def __init__(self, image, rank=None, world=None, device=None):
        import torch.distributed as dist
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.image = image if isinstance(image, _native.Image) else _native.Image.load(image)
        self.image.upload(device=device, shard_ix=self.rank, n_shards=self.world)

    def offtarget_search(self, protos, **kw):
        """Merged, canonically ordered hits of ALL shards on every rank (+ OR-ed key-presence flags)."""
        import torch
        import torch.distributed as dist
        hits, flags = self.image.offtarget_search(protos, **kw)
        merged = all_gather_hits(hits)
        if flags is not None:
            t = torch.from_numpy(flags.astype(np.int32))
            t = t.cuda() if dist.get_backend() == "nccl" else t
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            flags = t.cpu().numpy().astype(np.uint8)
        return merged, flags

    def pam_scan(self, pam, mode=_native.SCAN_GUIDE):
        fw, rv = self.image.pam_scan_genome(pam, mode)
        return all_gather_positions(fw), all_gather_positions(rv)
