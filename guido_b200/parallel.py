"""Multi-GPU data parallelism for the hot path: one process per GPU (torchrun).

The reference has no parallel code (its only concurrency is bowtie's ``-p``).  Here every window
of the genome is independent, so the padded genome coordinate is cut into ``world_size``
contiguous, 65536-base aligned chunks (``guido_shard_bounds``); rank r uploads chunk r plus a
256-base halo and owns the PAM starts inside it, so no window is reported twice and none is lost.
Guides are replicated.  The only exchange is the merge of the per-rank result lists:

* off-target hits: ONE padded all-gather of fixed-size slots whose first record carries the rank's
  count (``DeviceMerge``: search, gather, concat and sort stay on the GPUs; NCCL over NVLink), or,
  without any collective, copies into receive buffers the ranks of a node share through CUDA IPC
  (``PeerGather``: copy engines, next to a search that fills every SM); ``all_gather_hits`` is the
  host-array form (counts, then padded records) that also runs on gloo for the CPU tests.
  ``ShardedGenome(partition="key")`` splits the core-key space of the PAM-site table instead of the
  genome, so the per-search guide index is built in N parts instead of N times;
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


class PeerGather:
    """All-gather of fixed-size slots WITHOUT a collective: every rank owns a receive buffer of
    ``n_slots x world x slot_words`` 8-byte words that the other ranks of the node map through CUDA IPC
    (``guido_peer_alloc`` / ``guido_peer_open``), and ``send(slot, src)`` copies this rank's part straight
    into every rank's buffer with ``world`` device-to-device copies on a stream of its own.  Between GPUs
    those copies run on the copy engines over NVLink: they need no SM, so they proceed next to a kernel
    that occupies every SM (the off-target join does) -- an NCCL all-gather's kernels would queue
    behind it.  ``fence()`` orders the step: one 1-element all-reduce behind this rank's copies; when it
    completes on a rank, every rank's copies into that rank's buffer have completed.

    The layout is the one ``guido_packed_concat_dev`` reads: slot k, rank r at word
    ``(k * world + r) * slot_words``, header word first."""

    def __init__(self, device, n_slots, slot_words, group=None):
        import ctypes
        import torch
        import torch.distributed as dist
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.dev = torch.device("cuda", device) if isinstance(device, int) else device
        self.n_slots, self.slot_words = int(n_slots), int(slot_words)
        self.words = self.n_slots * self.world * self.slot_words
        lib = _native.lib()
        self._lib = lib
        # every rank takes part in every collective below whatever happens to it locally: a rank that
        # cannot allocate or map says so in the agreement at the end, and then ALL ranks raise
        ptr, handle = ctypes.c_void_p(), (ctypes.c_uint8 * 64)()
        err = None
        if lib.guido_peer_alloc(self.dev.index, self.words * 8, ctypes.byref(ptr), handle) != 0:
            err = _native.last_error()
        self.local = ptr.value
        mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=self.dev)
        handles = torch.empty((self.world, 64), dtype=torch.uint8, device=self.dev)
        dist.all_gather_into_tensor(handles, mine, group=group)
        good = torch.tensor([0 if err else 1], dtype=torch.int32, device=self.dev)
        dist.all_reduce(good, op=dist.ReduceOp.MIN, group=group)
        self.peers = [None] * self.world
        if int(good.item()):
            for r, h in enumerate(handles.cpu().numpy()):
                if r == self.rank:
                    self.peers[r] = self.local
                    continue
                p = ctypes.c_void_p()
                buf = (ctypes.c_uint8 * 64)(*h.tolist())
                if lib.guido_peer_open(self.dev.index, buf, ctypes.byref(p)) != 0:
                    err = _native.last_error()
                    break
                self.peers[r] = p.value
        good = torch.tensor([0 if err else 1], dtype=torch.int32, device=self.dev)
        dist.all_reduce(good, op=dist.ReduceOp.MIN, group=group)
        if not int(good.item()):
            self.close()
            raise RuntimeError(f"peer buffers are not available on this node: {err or 'another rank failed'}")
        self.stream = torch.cuda.Stream(device=self.dev)
        # one stream per destination: the copies to different ranks run on different copy engines
        self.lanes = [torch.cuda.Stream(device=self.dev) for _ in range(self.world)]
        self.lane_done = [torch.cuda.Event() for _ in range(self.world)]
        self.token = torch.zeros(1, dtype=torch.int32, device=self.dev)
        for ev, lane in zip(self.lane_done, self.lanes):
            ev.record(lane)   # (creates the handles)
        # raw handles for the one-call-per-slot C entry points
        self._c_peers = (ctypes.c_void_p * self.world)(*[ctypes.c_void_p(p) for p in self.peers])
        self._c_lanes = (ctypes.c_void_p * self.world)(*[ctypes.c_void_p(l.cuda_stream) for l in self.lanes])
        self._c_done = (ctypes.c_void_p * self.world)(*[ctypes.c_void_p(e.cuda_event) for e in self.lane_done])

    def send(self, slot, src_ptr, after=None):
        """this rank's ``slot_words`` words at ``src_ptr`` -> its place in every rank's buffer; the copies start
        behind the event ``after`` (the data is ready), one stream per destination"""
        at = (slot * self.world + self.rank) * self.slot_words * 8
        _native.check(self._lib.guido_peer_send(self._c_peers, self._c_lanes, self.world, self.rank, at, src_ptr,
                                                self.slot_words * 8, after.cuda_event if after is not None else None))

    def fence(self):
        """behind it (on self.stream), every rank's sends issued before ITS fence have landed everywhere"""
        import torch
        import torch.distributed as dist
        _native.check(self._lib.guido_streams_join(self._c_lanes, self._c_done, self.world, self.stream.cuda_stream))
        with torch.cuda.stream(self.stream):
            dist.all_reduce(self.token, group=self.group)

    def close(self):
        lib = self._lib
        for r, p in enumerate(self.peers):
            if r != self.rank and p:
                lib.guido_peer_close(p)
        if self.local:
            lib.guido_peer_free(self.local)
        self.peers, self.local = [], None


class DeviceMerge:
    """Off-target search of this rank's share + merge of all ranks' hits without leaving the GPUs
    (NCCL process group).  The letters go up once and are packed on the device, the hits stay in
    HBM (``guido_offtarget_search_dev``), and the per-rank lists travel in ONE fixed-size all-gather:
    every rank sends ``stride`` records, the first one a header carrying its hit count (written by a
    device-side copy), so no host round trip sits between search, gather and merge
    (``guido_hits_concat_dev``).  The merged list is brought into canonical (guide, gpos, strand)
    order on every GPU (``guido_hits_sort_dev``).

    ``to_host``:
      * ``"rank0"`` / True -- rank 0 copies the list to pinned host memory (16-byte records);
      * ``"shared"`` -- every rank packs the list to 8-byte records (``guido_hits_pack_dev``) and copies
        ITS 1/world of it into one host segment shared by the ranks of the node (a ``/dev/shm`` file,
        page-locked in every process): all PCIe links carry 1/world of the bytes instead of one link
        carrying all of them.  Returns the whole packed list (a view of the shared segment);
      * False -- the sorted (n, 4) int32 device tensor.
    Buffers are kept between calls; ``stride`` follows the largest per-rank count seen so far."""

    def __init__(self, image, group=None):
        import torch
        self.image, self.group = image, group
        self.dev = torch.device("cuda", int(image.info.device))  # the device the image was uploaded to
        self.stride = 0
        self.d_send = self.d_gath = self.d_merged = self.d_packed = None
        self.d_count = self.d_total = self.d_bad = self.host = None
        self.shared = None

    def _shared_segment(self, n_records):
        """uint64 view of a host segment all ranks of the node map and page-lock"""
        import mmap
        import os
        import torch
        import torch.distributed as dist
        need = max(int(n_records), 1) * 8
        if self.shared is not None and self.shared[0] >= need:
            return self.shared[1]
        rank = dist.get_rank(self.group)
        size = (need + need // 4 + (1 << 20)) & ~4095
        if self.shared is not None:
            torch.cuda.cudart().cudaHostUnregister(self.shared[3])
            self.shared[2].close()
        path = f"/dev/shm/guido_b200_merge_{os.environ.get('MASTER_PORT', '0')}_{os.getuid()}_{size}"
        if rank == 0:
            with open(path, "wb") as fh:
                fh.truncate(size)
        dist.barrier(group=self.group)
        fd = os.open(path, os.O_RDWR)
        mm = mmap.mmap(fd, size)
        os.close(fd)
        dist.barrier(group=self.group)
        if rank == 0:
            os.unlink(path)  # the mappings keep the segment alive
        arr = np.frombuffer(mm, dtype=np.uint64)
        rc = torch.cuda.cudart().cudaHostRegister(arr.ctypes.data, size, 0)
        if int(rc) != 0:
            raise RuntimeError(f"cudaHostRegister of the shared merge segment failed: {rc}")
        self.shared = (size, arr, mm, arr.ctypes.data)
        return arr

    def search(self, protos, pam="NGG", core_length=10, core_mismatches=2, total_mismatches=4,
               algo=_native.OT_AUTO, to_host=True):
        """-> merged hits, identical on every rank (see the class docstring for ``to_host``)"""
        import os
        import sys
        import time
        import torch
        import torch.distributed as dist
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        trace = os.environ.get("GUIDO_TRACE") and rank == 0
        t0 = time.perf_counter()

        def lap(what):  # GUIDO_TRACE=1: wall clock of the phases on stderr (rank 0; synchronises)
            if trace:
                torch.cuda.synchronize(self.dev)
                print(f"[guido merge] {what:18s} {(time.perf_counter() - t0) * 1e3:8.3f} ms", file=sys.stderr)

        if isinstance(protos, np.ndarray):
            letters = np.ascontiguousarray(protos, dtype=np.uint8).reshape(-1, 20)
        else:
            protos = list(protos)
            for p in protos:
                if len(p) != 20:
                    raise ValueError(f"protospacer must be 20 nt, got {len(p)}: {p!r}")
            letters = np.frombuffer("".join(protos).encode(), np.uint8).reshape(-1, 20)
        n = letters.shape[0]
        stream = torch.cuda.current_stream(self.dev)
        with torch.cuda.device(self.dev):
            # the letters go up as they are and are packed into plane words on the device
            d_letters = torch.from_numpy(letters.copy() if not letters.flags.writeable else letters).to(self.dev, non_blocking=True)
            d_guides = torch.empty((max(n, 1), 2), dtype=torch.int32, device=self.dev)
            if self.d_count is None:
                self.d_count = torch.zeros(1, dtype=torch.int64, device=self.dev)
                self.d_total = torch.zeros(1, dtype=torch.int64, device=self.dev)
                self.d_bad = torch.empty(1, dtype=torch.int64, device=self.dev)
            self.d_bad.fill_(-1)
            self.image.encode_guides_dev(d_letters.data_ptr(), n, d_guides.data_ptr(), self.d_bad.data_ptr(),
                                         stream.cuda_stream)
            if self.stride == 0:
                self.stride = max(1 << 16, 64 * n // world) + 1
            while True:
                if self.d_send is None or self.d_send.shape[0] < self.stride:
                    self.d_send = torch.empty((self.stride, 4), dtype=torch.int32, device=self.dev)
                    self.d_gath = torch.empty((world * self.stride, 4), dtype=torch.int32, device=self.dev)
                    self.d_merged = torch.empty((world * self.stride, 4), dtype=torch.int32, device=self.dev)
                stride = self.d_send.shape[0]
                self.image.offtarget_search_dev(d_guides.data_ptr(), n, pam, core_length, core_mismatches,
                                                total_mismatches, algo, self.d_send[1:].data_ptr(), stride - 1,
                                                self.d_count.data_ptr(), stream.cuda_stream)
                self.d_send[0, :2].copy_(self.d_count.view(torch.int32))  # the header: this rank's count
                lap("search")
                dist.all_gather_into_tensor(self.d_gath, self.d_send, group=self.group)
                lap("gather")
                self.image.concat_hits_dev(self.d_gath.data_ptr(), world, stride, self.d_merged.data_ptr(),
                                           self.d_merged.shape[0], self.d_total.data_ptr(), stream.cuda_stream)
                # the one host read of the call: every rank's count (overflow check), the total, bad letters
                heads = self.d_gath[::stride, :2].contiguous().view(torch.int64).reshape(-1)
                info = torch.cat([heads, self.d_total, self.d_bad]).tolist()
                counts, total, bad = info[:world], info[world], info[world + 1]
                lap("concat")
                if bad != -1:
                    raise ValueError(f"guide {bad // 32} holds a letter outside A/C/G/T at position {bad % 32 + 1}")
                if max(counts) <= stride - 1:
                    break
                self.stride = max(counts) + max(counts) // 8 + 1  # the exact need is known now: one more pass
                self.d_send = None
            merged = self.d_merged[:total]
            if total > 1:
                self.image.sort_hits_dev(merged.data_ptr(), total, n, stream.cuda_stream)
            lap("sort")
            if not to_host:
                return merged
            if to_host == "shared":
                if self.d_packed is None or self.d_packed.shape[0] < total:
                    self.d_packed = torch.empty(total + total // 8 + 16, dtype=torch.int64, device=self.dev)
                self.image.pack_hits_dev(merged.data_ptr(), total, self.d_packed.data_ptr(), stream.cuda_stream)
                seg = self._shared_segment(total)
                lo, hi = total * rank // world, total * (rank + 1) // world
                if hi > lo:
                    dst = torch.from_numpy(seg[lo:hi].view(np.int64))
                    dst.copy_(self.d_packed[lo:hi], non_blocking=True)
                stream.synchronize()
                dist.barrier(group=self.group)  # every part of the segment is in place
                lap("on the host")
                return seg[:total]
            if to_host == "rank0" and rank != 0:
                return merged
            if self.host is None or self.host.shape[0] < total:
                self.host = torch.empty((total + total // 8 + 16, 4), dtype=torch.int32, pin_memory=True)
            self.host[:total].copy_(merged, non_blocking=True)
            stream.synchronize()
            lap("on the host")
        return self.host[:total].numpy().view(_native.OT_HIT_DTYPE).reshape(-1)


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
        self._merge = None
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
        import torch.distributed as dist
        protos = list(protos) if not isinstance(protos, np.ndarray) else [bytes(p).decode() for p in protos]
        if dist.get_backend() == "nccl":  # search, merge and sort stay on the GPUs
            if self._merge is None:
                self._merge = DeviceMerge(self.image)
            merged = self._merge.search(protos, **kw).copy()
        else:
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
