"""Multi-GPU data parallelism for the hot path: one process per GPU (torchrun).

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
    """This rank's share of a genome image, resident on this rank's GPU.

    ``partition="position"`` (default): the genome is cut into contiguous shards (see above); right
    for the PAM scan and for any search.  ``partition="key"``: every rank uploads the WHOLE image
    and restricts its PAM-site table -- and with it the per-search guide index -- to 1/world of the
    core-key space (``guido_site_table_partition``; world must be a power of two): off-target
    searches then do no replicated work at all.  ``pam_scan`` is position based either way (with
    ``partition="key"`` every rank holds the whole genome and rank 0's scan is the answer).
    """

    def __init__(self, image, rank=None, world=None, device=None, partition="position"):
        import torch.distributed as dist
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        if partition not in ("position", "key"):
            raise ValueError(f"unknown partition {partition!r}")
        self.partition = partition
        self.image = image if isinstance(image, _native.Image) else _native.Image.load(image)
        if partition == "key":
            self.image.upload(device=device)
            self.image.site_table_partition(self.rank, self.world)
        else:
            self.image.upload(device=device, shard_ix=self.rank, n_shards=self.world)

    @staticmethod
    def _reduce_max(flags):
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(flags.astype(np.int32))
        t = t.cuda() if dist.get_backend() == "nccl" else t
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.cpu().numpy().astype(np.uint8)

    def offtarget_search(self, protos, want_raw_flags=True, **kw):
        """Merged, canonically ordered hits of ALL ranks on every rank (+ key-presence flags).

        Key presence (a guide has any raw bowtie alignment) needs a brute-force pass over every
        window for the guides without a valid hit.  Run per rank, that pass would hit nearly every
        guide -- most guides have no hit in a given shard or key range -- so it is done once,
        globally: ranks search without it, OR their "has a valid hit" flags, and only the guides
        that have none anywhere are checked, split over the ranks.
        """
        protos = list(protos) if not isinstance(protos, np.ndarray) else [bytes(p).decode() for p in protos]
        hits, _ = self.image.offtarget_search(protos, want_raw_flags=False, **kw)
        merged = all_gather_hits(hits)
        if not want_raw_flags:
            return merged, None
        flags = np.zeros(len(protos), np.uint8)
        flags[np.unique(merged["guide"])] = 1
        todo = np.nonzero(flags == 0)[0]
        if len(todo):
            rule = {k: kw[k] for k in ("pam", "core_length", "core_mismatches", "total_mismatches") if k in kw}
            extra = np.zeros(len(protos), np.uint8)
            if self.partition == "key":   # whole genome on every rank: split the guides
                mine = todo[self.rank::self.world]
            else:                         # a stretch of the genome per rank: every rank checks all of them
                mine = todo
            if len(mine):
                extra[mine] = self.image.offtarget_raw_flags([protos[i] for i in mine], **rule)
            flags |= self._reduce_max(extra)
        return merged, flags

    def pam_scan(self, pam, mode=_native.SCAN_GUIDE):
        fw, rv = self.image.pam_scan_genome(pam, mode)
        if self.partition == "key":  # every rank scanned the whole genome
            return fw, rv
        return all_gather_positions(fw), all_gather_positions(rv)
