#!/usr/bin/env python
"""bench.py -- guido hot path on B200: off-target guides/sec and PAM-scan GB/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload offtarget|pam|mmej|locus] [--genome human|agam|<Mb>] [--guides G]
                    [--algo auto|scan|index|table] [--shard key|position]

One "step" = one pass of the hot path over one batch:
  offtarget  (default) 100 000 SpCas9 guides, <=4 mismatches (core 10 / 2), against a synthetic
             3.1 Gb genome -- BASELINE.json configs[3]; `--genome agam --guides 10000` is configs[2].
             A step builds the guide-side seed index and joins it with the genome's PAM-site table
             (every PAM-valid window, grouped by core k-mer; guide independent, built with the genome
             image before the timed region the way bowtie-build precedes the reference's search;
             `--algo index` streams the genome instead and needs no table).
  pam        whole-genome NGG scan, both strands, synthetic 273 Mb genome -- configs[1].
             The default run also measures this one and reports it under "pam_scan".
  mmej       configs[4]: Cas12a TTTV scan of a synthetic 100 Mb genome, guides kept inside the
             exons of 5 000 genes, MMEJ enumeration for every guide (one step = scan + batch kernel).
  locus      configs[0] (the reference's own CPU-runnable case): find_guides + off-targets <=3 mm
             for a 100 kb locus of a 10 Mb genome, through the drop-in Python API.
Multi-GPU: one process per GPU (torchrun).  Off-target search: every rank holds the whole image and
1/N of the PAM-site table, split by core k-mer range, and builds the guide index for that range
only (`--shard key`, default); or the genome is sharded in contiguous, tile-aligned chunks with a
256-base halo and the guides are replicated (`--shard position`; always for the PAM scan).  The
per-rank hit lists are merged with ONE padded NCCL all-gather inside the step (header record =
count; no host round trip).  Total work is fixed as N grows ("strong").

Prints ONE JSON line (rank 0).  `value` is timed with CUDA events with inputs resident in HBM;
`e2e` goes through the host-buffer C-ABI call (guide letters H2D, search, device sort, raw-hit flags,
8-byte hit records D2H into pinned host memory -- all inside the region).

`--impl reference` (and `cpu_baseline`): the CPU arm.  The reference's off-target path is the
external bowtie 1.3.1 binary, which is not in this image; what runs is the oracle's SEED-INDEXED C
restatement of the same search (index built once = bowtie-build, not timed; every step searches a
bounded sample of the guides against the WHOLE genome on all host threads), labelled kind "port".
PAM scan / guide construction / MMEJ are timed on the reference's own Python code where
/root/reference is mounted (kind "reference-python"), on the oracle port elsewhere.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "off-target guides/sec (<=4 mm, 3.1 Gb genome); PAM-scan genome GB/s vs HBM peak"
AGAM_LENS = [49_400_000, 61_500_000, 42_000_000, 53_200_000, 24_400_000, 42_400_000, 240_000, 15_000]
_HUMAN_MB = [248.96, 242.19, 198.30, 190.21, 181.54, 170.81, 159.35, 145.14, 138.39, 133.80, 135.09,
             133.28, 114.36, 107.04, 101.99, 90.34, 83.26, 80.37, 58.62, 64.44, 46.71, 50.82, 156.04, 57.23]
HUMAN_LENS = [int(x * 3.1e9 / sum(_HUMAN_MB)) for x in _HUMAN_MB]
PAM = "NGG"
OT_KW = dict(core_length=10, core_mismatches=2, total_mismatches=4)


def genome_lens(spec):
    if spec == "human":
        return HUMAN_LENS, 1004, 0.01, "synthetic GRCh38-sized 3.1 Gb genome, 24 contigs, 1% N"
    if spec == "agam":
        return AGAM_LENS, 1002, 0.005, "synthetic AgamP4-sized 273 Mb genome, 8 contigs, 0.5% N"
    mb = float(spec)
    n = max(1, min(8, int(mb // 5) or 1))
    return [int(mb * 1e6 / n)] * n, 1010, 0.005, f"synthetic {mb:g} Mb genome, {n} contigs, 0.5% N"


def sample_guides(img, n, seed):
    """n protospacers taken from N-free NGG sites of the genome (forward strand), seeded."""
    rng = np.random.default_rng(seed)
    contigs = img.contigs
    lens = np.array([c[1] for c in contigs], dtype=np.int64)
    offs = np.array([c[2] for c in contigs], dtype=np.int64)
    out, have = [], 0
    while have < n:
        m = max(4096, (n - have) * 20)
        ci = rng.choice(len(contigs), size=m, p=lens / lens.sum())
        pos = (rng.random(m) * np.maximum(lens[ci] - 23, 1)).astype(np.int64)
        w = img.fetch_windows(offs[ci] + pos, 23)
        ok = (w[:, 21] == ord("G")) & (w[:, 22] == ord("G")) & ~(w == ord("N")).any(axis=1)
        out.append(w[ok][:, :20])
        have += int(ok.sum())
    return np.ascontiguousarray(np.concatenate(out)[:n])


def host_threads():
    """threads the CPU arm uses: every core this process may run on (torchrun exports
    OMP_NUM_THREADS=1, which is about its own workers, not about this measurement)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the bench runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def summary(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        inside = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.05]
        rows = inside or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "in_timed_region": bool(inside), "reasons": sorted(reasons)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


def ncu_traffic(kernel, workload_ok):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json;
    ncu cannot run inside a timed bench); only reported when this run is the workload it was taken on"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not workload_ok or not os.path.exists(p):
        return None
    return json.load(open(p)).get(kernel, {}).get("bytes")


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured copy bandwidth (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------ CPU arm

class CpuOfftarget:
    """Seed-indexed C port of the off-target search on the host cores (oracle/guido_oracle.c).

    The index over the whole genome is built once (the bowtie-build counterpart, reported but not
    timed); a step searches a bounded sample of the guides against it with every host thread."""

    def __init__(self, img, build_budget_s=150.0):
        from oracle import oracle as O
        self.O = O
        self.cores = O.set_threads(host_threads())
        t0 = time.perf_counter()
        contigs, bases = [], 0
        total = sum(c[1] for c in img.contigs)
        # contigs are added largest-first until the build budget would be overrun (a small host):
        # the windows-per-bucket, and with it the per-guide work, scale with the indexed fraction
        order = sorted(range(len(img.contigs)), key=lambda i: -img.contigs[i][1])
        est_rate = None
        for n_done, i in enumerate(order):
            name, clen, goff = img.contigs[i]
            if est_rate is not None and (bases + clen) / est_rate > build_budget_s and n_done >= 2:
                break
            contigs.append((name, img.fetch(i, 0, clen).encode(), goff))
            bases += clen
            if n_done == 0:  # calibrate on the first contig
                t1 = time.perf_counter()
                probe = O.SeedIndex(contigs, PAM, OT_KW["core_length"])
                probe.close()
                est_rate = clen / max(time.perf_counter() - t1, 1e-3) * 0.8
        self.fraction = bases / total
        self.index = O.SeedIndex(contigs, PAM, OT_KW["core_length"])
        self.build_s = time.perf_counter() - t0
        self.n_windows = self.index.n_windows
        del contigs

    def step(self, guides_u8):
        t0 = time.perf_counter()
        hits = self.index.search(guides_u8, OT_KW["core_mismatches"], OT_KW["total_mismatches"])
        dt = time.perf_counter() - t0
        return len(guides_u8) / dt, hits, dt

    def baseline(self, guides_u8, budget_s=12.0):
        n = min(len(guides_u8), 2000)
        rate, hits, dt = self.step(guides_u8[:n])
        n = int(min(len(guides_u8), max(n, rate * budget_s)))
        rate, hits, dt = self.step(guides_u8[:n])
        return self.describe(rate, n, hits, dt)

    def describe(self, rate, n, hits, dt):
        scale = ""
        value = rate
        if self.fraction < 0.999:
            # per-guide work = bucket look-ups + windows per bucket x checks; the second term grows with
            # the indexed fraction, so scaling the whole rate by it is pessimistic for the CPU by at
            # most the look-up share
            value = rate * self.fraction
            scale = (f"; index covers {self.fraction:.2f} of the genome (build budget), rate scaled by that fraction")
        return {"value": value, "unit": "guides/s", "cores": self.cores, "kind": "port",
                "sample": f"seed-indexed C port of the bowtie -n search + guido's PAM filter (oracle/guido_oracle.c; bowtie "
                          f"1.3.1 itself is not in this image): {n} of the guides against {self.n_windows} PAM-valid "
                          f"windows ({self.fraction:.2f} of the genome), {self.cores} host threads, {dt:.2f} s, "
                          f"{hits} hits; index build {self.build_s:.1f} s not timed (bowtie-build counterpart){scale}",
                "index_build_s": self.build_s, "genome_fraction": self.fraction}


def reference_python():
    """the reference package itself (unmodified, imported from /root/reference through stub modules for
    its absent third-party imports), or None where the reference is not mounted (the GPU box)"""
    try:
        from oracle import ref_loader
        if ref_loader.available():
            return ref_loader.load_reference()
    except Exception:
        pass
    return None


def cpu_pam_sample(img, budget_s=8.0):
    """PAM scan on the host: the reference's own `_find_guides_in_interval` (two regex passes + one
    Guide per hit, guido/locus.py:158-211) where it can be imported, the oracle's C scan elsewhere;
    one thread, as the reference runs.  Expressed in the same algorithmic bytes/base as the GPU line."""
    from oracle import oracle as O
    O.lib()
    ref = reference_python()
    if ref is not None:
        n = min(img.contigs[0][1], 200_000)
        seq = img.fetch(0, 0, n)
        loc = ref.Locus(sequence=seq, name="chr1", start=1)
        t0 = time.perf_counter()
        reps, hits = 0, 0
        while time.perf_counter() - t0 < budget_s / 2:
            hits = len(loc._find_guides_in_interval(seq, 1, PAM))
            reps += 1
        dt = time.perf_counter() - t0
        bases_per_s = n * reps / dt
        bytes_per_base = 0.25 + 4.0 * hits / n
        return {"value": bases_per_s * bytes_per_base / 1e9, "unit": "GB/s", "cores": 1, "kind": "reference-python",
                "sample": f"reference Locus._find_guides_in_interval (regex scan + Guide objects, guido/locus.py:158-211), "
                          f"1 thread: {n} bp x {reps} reps in {dt:.2f} s = {bases_per_s / 1e6:.3f} Mb/s, "
                          "expressed in the same algorithmic bytes/base", "bases_per_s": bases_per_s}
    n = min(img.contigs[0][1], 4_000_000)
    seq = img.fetch(0, 0, n)
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < budget_s / 2:
        fw, rv = O.pam_scan(seq, PAM)
        reps += 1
    dt = time.perf_counter() - t0
    bases_per_s = n * reps / dt
    bytes_per_base = 0.25 + 4.0 * (len(fw) + len(rv)) / n
    return {"value": bases_per_s * bytes_per_base / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
            "sample": f"oracle C scan only, no Guide construction (1 thread, as the reference runs): {n} bp x {reps} reps "
                      f"in {dt:.2f} s = {bases_per_s / 1e6:.1f} Mb/s, expressed in the same algorithmic bytes/base",
            "bases_per_s": bases_per_s}


# ------------------------------------------------------------------------------ GPU arm

class Runner:
    def __init__(self, args, N, torch, dist, rank, world, local_rank):
        self.a, self.N, self.torch, self.dist = args, N, torch, dist
        self.rank, self.world = rank, world
        self.dev = torch.device("cuda", local_rank)
        self.stream = torch.cuda.current_stream()
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB of L2
        self.sampler = ClockSampler(local_rank) if rank == 0 else None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allsum(self, x, op=None):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=op or self.dist.ReduceOp.SUM)
        return float(t.item())

    def allmax(self, x):
        return self.allsum(x, self.dist.ReduceOp.MAX if self.world > 1 else None)

    def timed(self, step, steps, warmup):
        """W warm-up steps, then K steps timed by CUDA events on the launching stream (L2 flushed
        between steps, outside the events), bracketed by barrier + synchronize; max over ranks."""
        torch = self.torch
        for _ in range(warmup):
            step()
        self.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t0 = time.time()
        launches = 0
        for s, e in ev:
            self.flush.zero_()
            s.record(self.stream)
            launches += step()
            e.record(self.stream)
        self.barrier()
        t1 = time.time()
        ms = sum(s.elapsed_time(e) for s, e in ev)
        ms_step = self.allmax(ms) / steps
        clocks = self.sampler.summary(t0, t1) if self.sampler else None
        return ms_step, launches, clocks

    def wall(self, fn, reps):
        for _ in range(3):  # warm-up: the first calls size the device / pinned result buffers
            fn()
        self.barrier()
        t0 = time.perf_counter()
        out = None
        for _ in range(reps):
            out = fn()
        self.barrier()
        ms = (time.perf_counter() - t0) * 1e3 / reps
        return self.allmax(ms), out


def run_pam(R, img, steps, warmup):
    """whole-shard NGG scan; returns a dict with value / roofline / e2e pieces"""
    N, torch, dist = R.N, R.torch, R.dist
    info = img.info
    span = info.shard_hi - info.shard_lo
    cap = span // 8 + (1 << 20)
    d_fw = torch.empty(cap, dtype=torch.int32, device=R.dev)
    d_rv = torch.empty(cap, dtype=torch.int32, device=R.dev)
    d_counts = torch.zeros(2, dtype=torch.int64, device=R.dev)

    def step():
        img.pam_scan_genome_dev(PAM, N.SCAN_GUIDE, d_fw.data_ptr(), cap, d_rv.data_ptr(), cap,
                                d_counts.data_ptr(), R.stream.cuda_stream)
        if R.world > 1:  # merge: the per-rank lists are already globally ordered by shard; gather counts
            allc = torch.empty(2 * R.world, dtype=torch.int64, device=R.dev)
            dist.all_gather_into_tensor(allc, d_counts)
        return launches_per_scan

    # kernels per scan as the library counts them: 1 = pam_onepass_kernel (look-back), 2 = count + fill
    st0 = N.Stats()
    img.pam_scan_genome(PAM, N.SCAN_GUIDE, stats=st0, cap_hint=cap, pinned=True)
    launches_per_scan = int(st0.n_launches)
    ms_step, launches, clocks = R.timed(step, steps, warmup)
    kt = img.kernel_times(steps)
    kernel_ms = R.allmax(float(np.mean(kt)))
    hits_rank = float(d_counts.sum().item())
    G = info.total_bases
    hits = R.allsum(hits_rank)
    alg_bytes_rank = 0.25 * span + 4.0 * hits_rank
    alg_bytes = R.allsum(alg_bytes_rank)

    def e2e():
        st = N.Stats()
        fw, rv = img.pam_scan_genome(PAM, N.SCAN_GUIDE, stats=st, cap_hint=cap, pinned=True)
        return st

    e2e_ms, st = R.wall(e2e, 3)
    return {"ms_per_step": ms_step, "launches": launches, "clocks": clocks, "hits": int(hits), "bases": int(G),
            "gbs": alg_bytes / (ms_step * 1e-3) / 1e9, "kernel_ms": kernel_ms,
            "kernel_gbs_per_gpu": alg_bytes_rank / (kernel_ms * 1e-3) / 1e9,
            "e2e_ms": e2e_ms, "e2e_gbs": alg_bytes / (e2e_ms * 1e-3) / 1e9,
            "kernel": "pam_onepass_kernel" if launches_per_scan == 1 else "pam_count_kernel + pam_fill_kernel",
            "h2d": int(st.bytes_h2d), "d2h": int(st.bytes_d2h)}


def hits_checksum(torch, d_packed, n):
    """order-independent 64-bit checksum of n 8-byte hit records (guide << 33 | position << 1 | strand) on
    the device: sum of the mixed words -- equal at every N iff the merged lists are equal as sets"""
    if n == 0:
        return 0
    w = d_packed[:n]
    x = (w ^ (w >> 29)) * -4658895280553007687
    x = (x ^ (x >> 31)) * 6364136223846793005
    return int(x.sum().item()) & 0xFFFFFFFFFFFFFFFF


def run_offtarget(R, img, guides_u8, algo, steps, warmup):
    N, torch, dist = R.N, R.torch, R.dist
    n_guides = len(guides_u8)
    enc = N.encode_guides([bytes(g).decode() for g in guides_u8])
    d_guides = torch.from_numpy(enc.view(np.int32).copy()).to(R.dev)
    # this rank's hit records (16 bytes each)
    cap = max(1 << 20, 256 * n_guides // R.world + (1 << 18))
    d_hits = torch.empty((cap, 4), dtype=torch.int32, device=R.dev)
    d_count = torch.zeros(1, dtype=torch.int64, device=R.dev)
    d_total = torch.zeros(1, dtype=torch.int64, device=R.dev)
    use_index = algo != N.OT_SCAN
    # genome-side index (PAM-site table): guide independent, part of the built genome like the
    # reference's bowtie-build output, so it is built before -- not inside -- the timed region
    table = None
    if algo in (N.OT_AUTO, N.OT_TABLE):
        n_sites, build_ms = img.site_table_build(PAM, OT_KW["core_length"])
        # 16 B per slot (position, planes, key: read for hits) + 5 B per slot of bit-sliced blocks (what
        # the join streams); buckets are padded to whole blocks of 64 slots
        table = {"sites_per_gpu": n_sites, "bytes_per_gpu": 21 * n_sites + 21 * 32 * (4 ** OT_KW["core_length"]) // R.world,
                 "build_ms": build_ms}
    # kernels per search: one host-API call reports its launches; the device sort (split + radix
    # passes + merge) belongs to that call only
    st0 = N.Stats()
    os.environ["GUIDO_SORT"] = "radix"  # this one call: the general sort, whose launch count is known below
    try:
        img.offtarget_search(guides_u8, pam=PAM, algo=algo, want_raw_flags=False, stats=st0, pinned=True, **OT_KW)
    finally:
        del os.environ["GUIDO_SORT"]
    guide_bits = max(1, int(np.ceil(np.log2(max(n_guides, 2)))))
    launches_per_search = int(st0.n_launches) - (2 + (33 + guide_bits + 7) // 8 if st0.bytes_d2h >= 16 + 2 * 16 else 0)

    def search():
        img.offtarget_search_dev(d_guides.data_ptr(), n_guides, PAM, OT_KW["core_length"],
                                 OT_KW["core_mismatches"], OT_KW["total_mismatches"], algo,
                                 d_hits.data_ptr(), cap, d_count.data_ptr(), R.stream.cuda_stream)

    stride = 0
    n_slices = 1
    merge_how = None
    if R.world > 1:
        # The merge: every rank searches its key range in n_slices slices; behind the join of a slice its
        # hits are packed to the 8-byte wire format behind a header word (count, read on the device), and a
        # second stream copies that slot straight into every rank's receive buffer (CUDA IPC peer buffers:
        # copy engines over NVLink, no SM needed) WHILE the next slice is searched -- an NCCL all-gather's
        # kernels would queue behind the join, which fills every SM.  One tiny all-reduce orders the step,
        # one concat makes the list contiguous.  No host round trip inside the step.  The slot per
        # (slice, rank) is sized from a first search (largest count + 10 %): a slot that is cut shows in
        # its header, the check after the timed region raises.  GUIDO_BENCH_MERGE=nccl: one padded NCCL
        # all-gather per slice instead of the peer copies.
        sliced_ok = algo in (N.OT_AUTO, N.OT_TABLE) and table is not None
        n_slices = int(os.environ.get("GUIDO_BENCH_SLICES", "2")) if sliced_ok else 1
        while n_slices > 1 and n_slices * R.world > 32:   # a slice is a whole number of 1/32s of the key space
            n_slices //= 2
        slice_cap = cap // n_slices
        d_counts = torch.zeros(n_slices, dtype=torch.int64, device=R.dev)
        events = [torch.cuda.Event() for _ in range(n_slices)]
        for e in events:
            e.record(R.stream)
        done = torch.cuda.Event()

        def search_slices(d_send=None):
            if sliced_ok:
                img.offtarget_search_sliced_dev(d_guides.data_ptr(), n_guides, PAM, OT_KW["core_length"], OT_KW["core_mismatches"],
                                                OT_KW["total_mismatches"], n_slices, d_hits.data_ptr(), slice_cap,
                                                d_counts.data_ptr(), [e.cuda_event for e in events], R.stream.cuda_stream,
                                                d_send.data_ptr() if d_send is not None else None, stride)
            else:
                search()
                d_counts.copy_(d_count)
                if d_send is not None:
                    img.pack_send_dev(d_hits.data_ptr(), d_count.data_ptr(), stride - 1, d_send.data_ptr(), R.stream.cuda_stream)
                events[0].record(R.stream)

        search_slices()
        stride = int(R.allmax(float(d_counts.max().item())) * 1.1) + 4096
        d_send = torch.empty((n_slices, stride), dtype=torch.int64, device=R.dev)
        d_merged = torch.empty(n_slices * R.world * stride, dtype=torch.int64, device=R.dev)
        pg = None
        if os.environ.get("GUIDO_BENCH_MERGE", "peer") != "nccl":
            from guido_b200.parallel import PeerGather
            try:
                pg = PeerGather(R.dev, n_slices, stride)   # (raises on every rank or on none)
            except RuntimeError as exc:  # no peer access between these GPUs: the collective does it
                if R.rank == 0:
                    print(f"{exc}; merging with NCCL all-gathers", file=sys.stderr)
        merge_how = "peer copies (CUDA IPC, copy engines)" if pg is not None else "NCCL all-gather"
        comm = pg.stream if pg is not None else torch.cuda.Stream(device=R.dev)
        d_gath = None if pg is not None else torch.empty((n_slices, R.world * stride), dtype=torch.int64, device=R.dev)
        gath_ptr = pg.local if pg is not None else d_gath.data_ptr()

    def step():
        if R.world == 1:
            search()
            return launches_per_search
        search_slices(d_send)
        if pg is not None:
            for k in range(n_slices):
                pg.send(k, d_send[k].data_ptr(), after=events[k])
            pg.fence()
            done.record(comm)
        else:
            with torch.cuda.stream(comm):
                for k in range(n_slices):
                    comm.wait_event(events[k])
                    dist.all_gather_into_tensor(d_gath[k], d_send[k])
                done.record(comm)
        R.stream.wait_event(done)
        img.concat_packed_dev(gath_ptr, n_slices * R.world, stride, d_merged.data_ptr(), d_merged.shape[0],
                              d_total.data_ptr(), R.stream.cuda_stream)
        return launches_per_search + n_slices + 1

    ms_step, launches, clocks = R.timed(step, steps, warmup)
    kt = img.kernel_times(steps)
    region_ms = R.allmax(float(np.mean(kt)))
    kernel_ms = region_ms
    if R.world > 1:
        per_slice = d_counts.cpu().tolist()
        n_rank = int(sum(per_slice))
        if max(per_slice) > min(stride - 1, slice_cap):
            raise RuntimeError(f"merge slot too small: {max(per_slice)} > {min(stride - 1, slice_cap)}")
    else:
        n_rank = int(d_count.item())
        if n_rank > cap:
            raise RuntimeError(f"hit buffer too small: {n_rank} > {cap}")
    if R.world > 1:
        hits = int(d_total.item())
        checksum = hits_checksum(torch, d_merged, hits)
    else:
        hits = n_rank
        d_packed = torch.empty(max(hits, 1), dtype=torch.int64, device=R.dev)
        img.pack_hits_dev(d_hits.data_ptr(), hits, d_packed.data_ptr(), R.stream.cuda_stream)
        checksum = hits_checksum(torch, d_packed, hits)
        del d_packed
    if table is not None:
        # roofline pass: in the timed steps above the join runs as one launch per key slice next
        # to the index build of the following slice, so its event pair brackets both.  The same
        # steps again with the overlap switched off time the join kernel alone (one launch over
        # the whole table); these steps are not part of `value`.
        os.environ["GUIDO_OT_OVERLAP"] = "0"
        try:
            for _ in range(2):
                search()
            torch.cuda.synchronize()
            for _ in range(steps):
                R.flush.zero_()  # cold L2, like the timed steps
                search()
            torch.cuda.synchronize()
            kt = img.kernel_times(steps)
            kernel_ms = R.allmax(float(np.mean(kt)))
        finally:
            del os.environ["GUIDO_OT_OVERLAP"]

    if R.world > 1 and pg is not None:
        R.barrier()   # nobody copies into a buffer that is being released
        pg.close()

    # end to end through the public API, host buffers in and out.  One GPU:
    # guido_offtarget_search_csr (letters up, search, device sort, key-presence flags, the hit list as CSR
    # by guide -- 4.1 bytes per hit -- into pinned host memory).  Several GPUs: parallel.DeviceMerge -- letters up, search,
    # one NCCL all-gather, device concat + sort, then every rank copies ITS 1/N of the packed list into a
    # host segment shared by the ranks (all PCIe links busy instead of one).
    merge = None
    if R.world > 1:
        from guido_b200.parallel import DeviceMerge
        merge = DeviceMerge(img)

    def e2e():
        if merge is not None:
            h = merge.search(guides_u8, pam=PAM, algo=algo, to_host="shared", **OT_KW)
            return 20 * n_guides, 8 * len(h) // R.world + 8 * (R.world + 2)
        st = N.Stats()
        img.offtarget_search_csr(guides_u8, pam=PAM, algo=algo, want_raw_flags=True, stats=st, pinned=True, **OT_KW)
        return int(st.bytes_h2d), int(st.bytes_d2h)

    e2e_ms, (h2d, d2h) = R.wall(e2e, 3)
    st = st0  # windows / index keys of this rank's search (same search as the timed steps)
    info = img.info
    span = info.shard_hi - info.shard_lo
    # algorithmic bytes of the search kernel on this rank: genome planes once, one 8-byte bucket
    # lookup per window, 4 bytes per index key compared, 16 bytes per hit written
    sites, visited = float(st.n_sites), float(st.n_pairs) if use_index else 0.0
    hits_rank = float(n_rank)
    blocks = False
    if table is not None:
        # join over the site table.  Block join (>= 8 keys per bucket): 5 B per table slot (the 320-byte
        # blocks of 64 windows, buckets padded to whole blocks: + 32 slots per bucket on average), 8 B per
        # index entry, 8 B of bucket offsets per bucket, and 40 B per hit (8-byte entry and 16-byte slot
        # read at flush time, 16-byte record written).  Popc walk (small guide sets): 16 B per slot,
        # 4 B per key, 4 B per bucket offset, 36 B per hit.
        from math import comb
        n_keys = float(n_guides) * sum(comb(OT_KW["core_length"], k) * 3 ** k for k in range(OT_KW["core_mismatches"] + 1))
        n_buckets = 4.0 ** OT_KW["core_length"]
        parts = R.world if getattr(R.a, "key_sharded", False) else 1  # key-sharded ranks build 1/N of the index
        slots = sites + 32.0 * n_buckets / parts
        blocks = n_keys >= 8.0 * n_buckets and sites * parts >= 64.0 * n_buckets
        if blocks:
            alg_bytes_rank = 5.0 * slots + (8.0 * n_keys + 8.0 * n_buckets) / parts + 40.0 * hits_rank
        else:
            alg_bytes_rank = 16.0 * slots + (4.0 * n_keys + 4.0 * n_buckets) / parts + 36.0 * hits_rank
    else:
        alg_bytes_rank = 0.25 * span + (8.0 * sites + 4.0 * visited if use_index else 0.0) + 16.0 * hits_rank
    return {"ms_per_step": ms_step, "launches": launches, "clocks": clocks, "hits": hits, "checksum": checksum, "table": table,
            "kernel_ms": kernel_ms, "region_ms": region_ms, "kernel_gbs_per_gpu": alg_bytes_rank / (kernel_ms * 1e-3) / 1e9,
            "alg_bytes_rank": alg_bytes_rank, "sites": int(R.allsum(sites)), "visited": int(R.allsum(visited)),
            "pairs_scan": float(st.n_pairs) if not use_index else None, "blocks": blocks, "merge": merge_how, "n_slices": n_slices,
            "e2e_ms": e2e_ms, "h2d": h2d, "d2h": d2h}


# ------------------------------------------------------------------------------ configs[4]: TTTV + MMEJ

def exon_set(lens, seed=1006, n_genes=5000):
    """5 000 genes x 3-8 exons x 100-500 bp on a synthetic genome: (contig,
    start0, end0) per exon, non-overlapping inside a gene"""
    rng = np.random.default_rng(seed)
    lens = np.asarray(lens, dtype=np.int64)
    exons = []
    for _ in range(n_genes):
        ci = int(rng.choice(len(lens), p=lens / lens.sum()))
        at = int(rng.integers(1000, lens[ci] - 20000))
        for _ in range(int(rng.integers(3, 9))):
            ln = int(rng.integers(100, 501))
            exons.append((ci, at, at + ln))
            at += ln + int(rng.integers(100, 1500))
    return np.array(exons, dtype=np.int64)


def window_starts_in_exons(fw, rv, P, ex_lo, ex_end, run_lo, run_hi, flank=75):
    """Start of the 2 x flank window around every cut that lies in an exon and whose window is N-free; PAM
    positions fw / rv ascending (uint32), '+' sites first.

    The reference's geometry (guides.py:53-67): '+' cut = p - 3, '-' cut = p + P + 3.  A cut belongs to
    the last exon that starts at or before it, if it lies inside it: with the exons sorted these are the
    disjoint intervals [ex_lo[k], ex_end[k]) (ex_end = min(exon end, next exon start)), and the PAM
    positions whose cut falls inside are index ranges found by 2 x n_exons binary searches -- the host
    touches the ~8 % of the sites it keeps, not all of them.  Windows that touch a non-ACGT run are left
    out (the CPU arm skips guides with "N" in locus_seq): cut - flank < run end and cut + flank > run
    start, i.e. run_lo = run start - flank < cut < run end + flank = run_hi."""
    parts = []
    for pos, shift in ((fw, -3), (rv, P + 3)):
        first = np.searchsorted(pos, (ex_lo - shift).astype(np.uint32), side="left")
        last = np.searchsorted(pos, (ex_end - shift).astype(np.uint32), side="left")
        n_in = last - first
        ix = np.repeat(first - (np.cumsum(n_in) - n_in), n_in) + np.arange(int(n_in.sum()))
        cut = pos[ix].astype(np.int64) + shift
        r = np.searchsorted(run_hi, cut, side="right")           # first run whose reach ends behind the cut
        cut = cut[(r >= len(run_lo)) | (run_lo[np.minimum(r, len(run_lo) - 1)] >= cut)]
        parts.append(cut - flank)
    return np.concatenate(parts)


def run_mmej(R, args):
    """TTTV scan of a 100 Mb genome + MMEJ enumeration for every guide whose cut lies in an exon"""
    N, torch = R.N, R.torch
    from guido_b200 import _native
    lens = [12_500_000] * 8
    img = N.Image.synth(lens, seed=1006, n_frac=0.005).upload(device=R.dev.index)
    exons = exon_set(lens)
    pam, P = "TTTV", 4
    goff = np.array([c[2] for c in img.contigs], dtype=np.int64)
    ex_lo, ex_hi = goff[exons[:, 0]] + exons[:, 1], goff[exons[:, 0]] + exons[:, 2]
    order = np.argsort(ex_lo)
    ex_lo, ex_hi = ex_lo[order], ex_hi[order]

    run_st, run_en, _ = img.runs()
    ex_end = np.minimum(ex_hi, np.append(ex_lo[1:], np.iinfo(np.int64).max))
    run_lo, run_hi = run_st.astype(np.int64) - 75, run_en.astype(np.int64) + 75  # cuts in (run_lo, run_hi) are too close to the run

    laps = {"scan_ms": 0.0, "filter_ms": 0.0, "mmej_ms": 0.0}

    def step():
        t_a = time.perf_counter()
        fw, rv = img.pam_scan_genome(pam, N.SCAN_GUIDE, pinned=True)
        t_b = time.perf_counter()
        lo = window_starts_in_exons(fw, rv, P, ex_lo, ex_end, run_lo, run_hi)
        n_sites = len(fw) + len(rv)
        st = N.Stats()
        t_c = time.perf_counter()
        tuples, offsets, n_cand = img.mmej_sites(lo, 150, 75, stats=st, cap_hint=step.cap, pinned=True)
        t_d = time.perf_counter()
        laps["scan_ms"] += (t_b - t_a) * 1e3; laps["filter_ms"] += (t_c - t_b) * 1e3; laps["mmej_ms"] += (t_d - t_c) * 1e3
        step.cap = max(step.cap, int(len(tuples) * 1.05))
        return len(lo), int(n_cand.sum()), len(tuples), float(st.kernel_ms), n_sites

    step.cap = 0
    for _ in range(max(args.warmup, 1)):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    k_ms = 0.0
    steps = max(min(args.steps, 10), 1)
    for key in laps:
        laps[key] = 0.0
    for _ in range(steps):
        n_guides, n_cand, n_kept, kms, n_sites = step()
        k_ms += kms
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    k_ms /= steps
    out = {"metric": "guides/sec, Cas12a TTTV scan + MMEJ enumeration over a 5k-gene exon set (BASELINE configs[4])",
           "value": n_guides / (ms * 1e-3), "unit": "guides/s", "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 1),
           "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8 letters / u32 bit masks",
           "data": "synthetic",
           "config": {"workload": "TTTV scan of a synthetic 100 Mb genome (8 contigs), guides with the cut inside one of "
                                  f"{len(exons)} exons of 5000 genes, MMEJ candidates of every 150-bp window",
                      "pam_sites": n_sites, "guides": n_guides, "mmej_candidates": n_cand, "mmej_kept": n_kept,
                      "step_breakdown_ms": {k2: v / steps for k2, v in laps.items()},
                      "timing": "host wall clock of the whole step through the C-ABI calls (guido_pam_scan_genome -> exon / N-run "
                                "filter on the host -> guido_mmej_sites: windows read from the resident image, tuples to the host)"},
           "gpu_launches": steps * 2,
           "mmej_kernel": {"kernel": "mmej_fast_kernel", "kernel_ms": k_ms, "windows_per_s": n_guides / (k_ms * 1e-3),
                           "candidates_per_s": n_cand / (k_ms * 1e-3)},
           "e2e": {"value": n_guides / (ms * 1e-3), "unit": "guides/s", "ms_per_step": ms,
                   "h2d_bytes_per_step": 12 * n_guides, "d2h_bytes_per_step": 4 * n_sites + 4 * n_kept + 12 * n_guides}}
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_mmej_sample(img, exons, pam)
    return out


def cpu_mmej_sample(img, exons, pam, budget_s=15.0):
    """`_find_guides_in_interval(..., "TTTV")` + `generate_mmej_patterns` per guide on a bounded sample
    of the exons: the reference's own code where it is mounted, the oracle port elsewhere; 1 thread"""
    from oracle import oracle as O
    O.lib()
    ref = reference_python()
    t0 = time.perf_counter()
    n_guides = n_exons = 0
    for ci, a, b in exons.tolist():
        lo = max(a - 100, 0)
        seq = img.fetch(ci, lo, b - a + 200)
        if ref is not None:
            loc = ref.Locus(sequence=seq, name="c", start=lo + 1)
            for g in loc._find_guides_in_interval(seq, lo + 1, pam):
                if a < g.absolute_cut_pos <= b and "N" not in g.locus_seq:
                    ref.mmej.generate_mmej_patterns(g.relative_cut_pos, g.locus_seq, 20)
                    n_guides += 1
        else:
            for g in O.guide_records(seq, lo + 1, pam, chromosome="c"):
                if a < g["absolute_cut_pos"] <= b and "N" not in g["locus_seq"]:
                    O.mmej_tuples(g["locus_seq"], g["relative_cut_pos"])
                    n_guides += 1
        n_exons += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    kind = "reference-python" if ref is not None else "port"
    what = ("reference Locus._find_guides_in_interval + mmej.generate_mmej_patterns (guido/locus.py:158-211, mmej.py:8-120)"
            if ref is not None else "oracle port (Python guide records + C MMEJ enumeration, no pattern dicts)")
    return {"value": n_guides / dt, "unit": "guides/s", "cores": 1, "kind": kind,
            "sample": f"{what}, 1 thread: {n_exons} exons, {n_guides} guides in {dt:.1f} s"}


# ------------------------------------------------------------------------------ configs[0]: locus through the API

def run_locus(R, args):
    """find_guides + off-targets (<=3 mm) for chr2:1,000,001-1,100,000 of a synthetic 10 Mb genome through
    the drop-in Python API (Genome.build -> locus_from_coordinates -> find_guides -> find_off_targets)"""
    import tempfile
    import guido_b200 as guido
    N = R.N
    lens = [2_500_000] * 4
    names = ["chr1", "chr2", "chr3", "chr4"]
    img = N.Image.synth(lens, seed=1001, n_frac=0.005, names=names)
    tmp = tempfile.mkdtemp(prefix="guido_bench_")
    fa = os.path.join(tmp, "g10.fa")
    with open(fa, "w") as fh:
        for i, (name, clen, _) in enumerate(img.contigs):
            s = img.fetch(i, 0, clen)
            fh.write(f">{name}\n")
            fh.writelines(s[j:j + 60] + "\n" for j in range(0, len(s), 60))
    import contextlib
    t0 = time.perf_counter()
    g = guido.Genome("g10", genome_file_abspath=fa)
    with contextlib.redirect_stdout(sys.stderr):  # build() reports its progress on stdout, like the reference's
        g.build()
    build_s = time.perf_counter() - t0

    def once():
        loc = guido.locus_from_coordinates(g, "chr2", 1_000_001, 1_100_000)
        t0 = time.perf_counter()
        guides = loc.find_guides()
        t1 = time.perf_counter()
        found = loc.find_off_targets(total_mismatches=3)
        t2 = time.perf_counter()
        return len(guides), sum(len(v) for v in found.values()), (t1 - t0) * 1e3, (t2 - t1) * 1e3

    for _ in range(max(min(args.warmup, 3), 1)):
        once()
    fg, ot = [], []
    steps = max(min(args.steps, 5), 1)
    for _ in range(steps):
        n_guides, n_hits, a, b = once()
        fg.append(a); ot.append(b)
    ms = float(np.mean(fg) + np.mean(ot))
    out = {"metric": "guides/sec through the Python API: find_guides + off-target search (<=3 mm) on a 100 kb locus of a "
                     "10 Mb genome (BASELINE configs[0])",
           "value": n_guides / (ms * 1e-3), "unit": "guides/s", "n_gpus": 1, "steps": steps,
           "warmup": max(min(args.warmup, 3), 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "u32 bit planes", "data": "synthetic",
           "config": {"workload": "Genome.build (FASTA -> image) once; per step locus_from_coordinates(chr2:1000001-1100000) -> "
                                  "find_guides() -> find_off_targets(total_mismatches=3); host wall clock, Python records included",
                      "guides": n_guides, "off_target_records": n_hits, "find_guides_ms": float(np.mean(fg)),
                      "find_off_targets_ms": float(np.mean(ot)), "genome_build_s": build_s},
           "gpu_launches": None,
           "e2e": {"value": n_guides / (ms * 1e-3), "unit": "guides/s", "ms_per_step": ms,
                   "h2d_bytes_per_step": 20 * n_guides, "d2h_bytes_per_step": 16 * n_hits}}
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_locus_sample(img, names)
    return out


def cpu_locus_sample(img, names):
    """the reference's own find_guides + run_bowtie (its real parsing / filtering / record building, the bowtie
    subprocess replaced by the oracle's brute-force C search: oracle/fake_bowtie.py) on the same locus"""
    from oracle import oracle as O
    O.set_threads(host_threads())
    ref = reference_python()
    contigs = [(n, img.fetch(i, 0, img.contigs[i][1])) for i, n in enumerate(names)]
    seq = contigs[1][1][1_000_000:1_100_000]
    if ref is None:
        t0 = time.perf_counter()
        guides = O.find_guides_records(seq, 1_000_001, None, chromosome="chr2")
        t1 = time.perf_counter()
        sub = guides[:512]
        O.off_target_records(sub, contigs, total_mismatches=3)
        t2 = time.perf_counter()
        ot_s = (t2 - t1) * len(guides) / max(len(sub), 1)
        return {"value": len(guides) / ((t1 - t0) + ot_s), "unit": "guides/s", "cores": O.threads(), "kind": "port",
                "sample": f"oracle port: find_guides_records {len(guides)} guides in {t1 - t0:.2f} s (1 thread, Python); "
                          f"off_target_records of {len(sub)} guides (brute-force C search, {O.threads()} threads) in "
                          f"{t2 - t1:.2f} s, scaled to all guides"}
    from oracle import fake_bowtie
    loc = ref.Locus(sequence=seq, name="chr2", start=1_000_001)
    t0 = time.perf_counter()
    guides = loc.find_guides()
    t1 = time.perf_counter()
    loc.guides = guides[:512]
    loc.genome = fake_bowtie.FakeGenome("idx")
    with fake_bowtie.patched(ref, {"idx": contigs}):
        found = loc.find_off_targets(total_mismatches=3)
    t2 = time.perf_counter()
    ot_s = (t2 - t1) * len(guides) / 512
    return {"value": len(guides) / ((t1 - t0) + ot_s), "unit": "guides/s", "cores": O.threads(), "kind": "reference-python",
            "sample": f"reference Locus.find_guides: {len(guides)} guides in {t1 - t0:.2f} s (1 thread); reference "
                      f"find_off_targets/run_bowtie for 512 of them with the bowtie subprocess replaced by the oracle's brute-force "
                      f"C search ({O.threads()} threads): {t2 - t1:.2f} s, {sum(len(v) for v in found.values())} records, scaled to all guides"}


# ------------------------------------------------------------------------------ main

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="offtarget", choices=["offtarget", "pam", "mmej", "locus"])
    ap.add_argument("--genome", default=None)
    ap.add_argument("--guides", type=int, default=None)
    ap.add_argument("--algo", default="auto", choices=["auto", "scan", "index", "table"])
    ap.add_argument("--shard", default="key", choices=["key", "position"],
                    help="multi-GPU off-target search: split the core-key space (default) or the genome")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pam", action="store_true", help="skip the secondary PAM-scan measurement")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.genome is None:
        args.genome = "human" if args.workload == "offtarget" else "agam"
    if args.guides is None:
        args.guides = 100_000 if args.genome == "human" else 10_000
    lens, seed, n_frac, gdesc = genome_lens(args.genome)

    from guido_b200 import _native as N

    if args.impl == "reference":
        if rank == 0:
            run_reference(args, N, lens, seed, n_frac, gdesc)
        return 0

    args.warmup = max(args.warmup, 3)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # NCCL's version banner goes to stdout, which carries the ONE JSON line
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    algo = {"auto": N.OT_AUTO, "scan": N.OT_SCAN, "index": N.OT_INDEX, "table": N.OT_TABLE}[args.algo]
    R = Runner(args, N, torch, dist, rank, world, local_rank)
    peak, peak_src = measured_peak_gbs()

    if args.workload in ("mmej", "locus"):
        if rank == 0:
            out = run_mmej(R, args) if args.workload == "mmej" else run_locus(R, args)
            print(json.dumps(out), flush=True)
        if R.sampler:
            R.sampler.stop()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    img = N.Image.synth(lens, seed=seed, n_frac=n_frac)
    # Multi-GPU off-target search: by default every rank keeps the whole image (0.8 GB) and tables /
    # searches 1/N of the core-key space (guido_site_table_partition), so the per-search guide index
    # is built in N parts instead of N times.  `--shard position` gives every rank a stretch of
    # the genome instead (north_star's layout; what the PAM scan always uses).
    key_sharded = (args.workload == "offtarget" and world > 1 and args.shard == "key"
                   and algo in (N.OT_AUTO, N.OT_TABLE) and world & (world - 1) == 0 and world <= 32)
    if key_sharded:
        img.upload(device=local_rank).site_table_partition(rank, world)
    else:
        img.upload(device=local_rank, shard_ix=rank, n_shards=world)
    args.key_sharded = key_sharded
    out = {"metric": METRIC, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "data": "synthetic"}

    if args.workload == "pam":
        r = run_pam(R, img, args.steps, args.warmup)
        out.update({
            "value": r["gbs"], "unit": "GB/s", "ms_per_step": r["ms_per_step"], "dtype": "u32 bit-plane words",
            "gpu_launches": r["launches"], "clocks": r["clocks"],
            "config": {"workload": f"whole-genome NGG PAM scan, both strands, {gdesc} (BASELINE configs[1])",
                       "pam": PAM, "scan_mode": "guide windows (N-free, inside one contig)",
                       "l2": "flushed between steps (256 MB memset outside the events)",
                       "parallelism": f"genome sharded x{world}", "hits": r["hits"]},
            "bases_per_s": r["bases"] / (r["ms_per_step"] * 1e-3),
            "e2e": {"value": r["e2e_gbs"], "unit": "GB/s", "ms_per_step": r["e2e_ms"],
                    "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
            "roofline": {"bound": "hbm", "achieved": r["kernel_gbs_per_gpu"], "peak": peak, "unit": "GB/s",
                         "frac": r["kernel_gbs_per_gpu"] / peak,
                         "traffic": ncu_traffic(r["kernel"], args.genome == "agam" and world == 1),
                         "peak_source": peak_src,
                         "kernel": r["kernel"], "kernel_ms": r["kernel_ms"],
                         "algorithmic_bytes": "0.25 B/base read + 4 B/hit written, per GPU"},
        })
        if not args.no_cpu_baseline and world == 1 and rank == 0:
            out["cpu_baseline"] = cpu_pam_sample(img)
    else:
        guides_u8 = sample_guides(img, args.guides, seed + 1)
        r = run_offtarget(R, img, guides_u8, algo, args.steps, args.warmup)
        gps = args.guides / (r["ms_per_step"] * 1e-3)
        use_index = algo != N.OT_SCAN
        use_table = r["table"] is not None
        blocks = r["blocks"]
        kernel = ("ot_join_blocks_kernel" if blocks else "ot_join_kernel" if use_table else
                  "ot_kernel<index>" if use_index else "ot_kernel<scan>")
        algo_desc = ("guide-side seed index joined with the genome's PAM-site table" if use_table else
                     "seed index + streaming window scan" if use_index else "brute-force window x guide scan")
        alg_desc = ("per GPU: 5 B per slot of the bit-sliced site table (buckets padded to 64 slots) + 8 B per index "
                    "entry + 8 B of offsets per bucket + 40 B per hit; the compare itself is bound by the ALU pipe and the "
                    "shared-memory data pipe, see int_pipe" if blocks else
                    "per GPU: 16 B per table slot + 4 B per index key + 4 B per bucket offset + 36 B per hit" if use_table else
                    "per GPU: 0.25 B/base planes + 8 B bucket lookup per window + 4 B per index "
                    "key compared + 16 B per hit" if use_index else
                    "per GPU: 0.25 B/base planes + 16 B per hit (the brute-force compare is "
                    "integer-issue bound, see int_pipe)")
        out.update({
            "value": gps, "unit": "guides/s", "ms_per_step": r["ms_per_step"],
            "dtype": "u32 bit planes (bit-sliced LOP3 carry-save count)" if blocks else "u32 bit planes (XOR/OR + popc)",
            "gpu_launches": r["launches"], "clocks": r["clocks"],
            "config": {"workload": f"{args.guides} SpCas9 guides, <=4 mm (core 10 / <=2) vs {gdesc}"
                                   + (" (BASELINE configs[3] on N GPUs)" if args.genome == "human" else " (BASELINE configs[2])"),
                       "pam": PAM, "algo": algo_desc,
                       "l2": "flushed between steps; the shard's site table / planes + index keys exceed L2",
                       "parallelism": (f"PAM-site table partitioned by core k-mer range x{world} (whole image on every "
                                       "rank), guide index built in the same parts, one padded NCCL all-gather of the hit lists "
                                       "(count in a header record) + device concat"
                                       if key_sharded else
                                       f"genome sharded x{world}, guides replicated, one padded NCCL all-gather of the hit lists"),
                       "merge": (f"{r['n_slices']} key slices per rank; {r['merge']}" if world > 1 else None),
                       "hits": r["hits"], "hits_checksum": f"{r['checksum']:016x}",
                       "windows": r["sites"], "index_keys_compared": r["visited"],
                       "site_table": r["table"]},
            "e2e": {"value": args.guides / (r["e2e_ms"] * 1e-3), "unit": "guides/s", "ms_per_step": r["e2e_ms"],
                    "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                    "what": ("guido_offtarget_search_csr: 20 letters per guide up; search, device sort, key-presence flags; the hit "
                             "list as CSR by guide (4-byte position + 1 strand bit per hit, 4-byte offset + 1 flag byte per guide) "
                             "down into pinned host memory" if world == 1 else
                             "parallel.DeviceMerge: letters up on every rank; search, one all-gather, concat + sort on every GPU; "
                             "every rank copies its 1/N of the 8-byte records into a host segment shared by the ranks")},
            "roofline": {"bound": "hbm", "achieved": r["kernel_gbs_per_gpu"], "peak": peak, "unit": "GB/s",
                         "frac": r["kernel_gbs_per_gpu"] / peak,
                         "traffic": ncu_traffic(kernel, use_index and args.genome == "human" and args.guides == 100_000 and world == 1),
                         "algorithmic_bytes_per_launch": r["alg_bytes_rank"], "peak_source": peak_src,
                         "kernel": kernel, "kernel_ms": r["kernel_ms"], "algorithmic_bytes": alg_desc},
        })
        if use_table:
            out["roofline"]["kernel_ms_note"] = (
                "join kernel timed alone (CUDA events on its stream, one launch over the whole table, "
                "GUIDO_OT_OVERLAP=0, same steps re-run after the timed region); inside the timed steps the "
                "join runs as one launch per key slice overlapped with the index build of the next slice: "
                "first join launch .. last join done = join_region_ms")
            out["roofline"]["join_region_ms"] = r["region_ms"]
        if use_index and r["visited"]:
            sm_mhz = (r["clocks"] or {}).get("sm_mhz") or 1965.0
            per_gpu = r["visited"] / world / (r["kernel_ms"] * 1e-3)
            if blocks:
                # per (entry, 32 windows): one 4-byte shared-memory word per distal base (ten 8-byte loads per 64
                # windows) + 14 LOP3 (bias class "at most 2 of 10", most passes) or 18 LOP3 of carry-save counting on
                # the ALU pipe (16 lanes/clk per SM sub-partition = 2 warp-instr/clk/SM); `visited` = entries x slots
                out["int_pipe"] = {"pair_compares_per_s_per_gpu": per_gpu, "lop3_per_32_windows": "14 (bias 5 passes) / 18",
                                   "shared_loads_per_64_windows": 10,
                                   "alu_warp_instr_per_clk_per_sm_for_compare_at_14": per_gpu / 32 * 14 / 32 / 148 / (sm_mhz * 1e6),
                                   "alu_peak_warp_instr_per_clk_per_sm": 2, "sm_mhz": sm_mhz,
                                   "note": "no XOR/OR stage and no POPC: per-letter mismatch words are loaded, a LOP3 carry-save "
                                           "tree counts them; ncu (profiles/r2_join.md): ALU pipe and shared-memory data pipe both "
                                           "~65 % busy"}
            else:
                # the popc kernels do one POPC per key compared: 16 lanes/clk/SM on sm_100
                out["int_pipe"] = {"key_compares_per_s_per_gpu": per_gpu, "popc_per_clk_per_sm": per_gpu / 148 / (sm_mhz * 1e6),
                                   "popc_peak_per_clk_per_sm": 16, "sm_mhz": sm_mhz,
                                   "note": "XOR, fold (SHF+LOP3), POPC, ISETP per key; padding keys of the 8-key sectors not counted"}
        if not use_index and r["pairs_scan"]:
            out["int_pipe"] = {"pairs_per_s_per_gpu": r["pairs_scan"] / (r["kernel_ms"] * 1e-3), "ops_per_pair": 4,
                               "note": "2 LOP3 + POPC + ISETP per (window, guide) pair"}
        if not args.no_cpu_baseline and world == 1 and rank == 0:
            # the CPU arm of the same workload, beside the GPU number (index over the whole genome, a
            # bounded sample of the guides per measurement)
            img.site_table_free()
            cpu = CpuOfftarget(img)
            out["cpu_baseline"] = cpu.baseline(guides_u8)
            cpu.index.close()
        if not args.no_pam and world == 1:
            # second half of the metric in the same run: NGG scan of the 273 Mb genome (configs[1])
            img.close()
            lens2, seed2, nf2, gdesc2 = genome_lens("agam")
            img2 = N.Image.synth(lens2, seed=seed2, n_frac=nf2).upload(device=local_rank)
            p = run_pam(R, img2, 10, 3)
            out["pam_scan"] = {"value": p["gbs"], "unit": "GB/s", "ms_per_step": p["ms_per_step"],
                               "workload": f"whole-genome NGG PAM scan, both strands, {gdesc2} (BASELINE configs[1])",
                               "hits": p["hits"], "bases_per_s": p["bases"] / (p["ms_per_step"] * 1e-3),
                               "roofline": {"bound": "hbm", "achieved": p["kernel_gbs_per_gpu"], "peak": peak, "unit": "GB/s",
                                            "frac": p["kernel_gbs_per_gpu"] / peak, "kernel": p["kernel"],
                                            "kernel_ms": p["kernel_ms"],
                                            "traffic": ncu_traffic(p["kernel"], True)},
                               "e2e": {"value": p["e2e_gbs"], "unit": "GB/s", "ms_per_step": p["e2e_ms"],
                                       "d2h_bytes_per_step": p["d2h"]}}
            if not args.no_cpu_baseline and rank == 0:
                out["pam_scan"]["cpu_baseline"] = cpu_pam_sample(img2)
    if R.sampler:
        R.sampler.stop()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_reference(args, N, lens, seed, n_frac, gdesc):
    """CPU arm.  Off-target search: the seed-indexed C port on every host thread, index over the whole
    genome built once (bowtie-build counterpart), exactly --steps timed searches of a bounded guide
    sample after --warmup untimed ones.  PAM scan: the reference's own Python where it is mounted."""
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    if args.workload in ("mmej", "locus"):
        print(json.dumps({"impl": "reference", "unavailable": f"the CPU arm of --workload {args.workload} is the cpu_baseline "
                                                                "object of the b200 line"}), flush=True)
        return
    img = N.Image.synth(lens, seed=seed, n_frac=n_frac)
    G = img.info.total_bases
    vals = []
    if args.workload == "pam":
        cb = None
        for i in range(warmup + steps):
            cb = cpu_pam_sample(img, budget_s=max(1.0, min(6.0, 120.0 / (warmup + steps))))
            if i >= warmup:
                vals.append(cb["value"])
        v = float(np.mean(vals))
        cb["value"] = v
        ms = G / cb["bases_per_s"] * 1e3
        workload = f"whole-genome NGG PAM scan, both strands, {gdesc}"
        dtype = "regex over str" if cb["kind"] == "reference-python" else "u8 letters"
    else:
        guides_u8 = sample_guides(img, args.guides, seed + 1)
        cpu = CpuOfftarget(img)
        # sample size per step: about 100 s of search in total
        rate, _, _ = cpu.step(guides_u8[:1000])
        n = int(min(len(guides_u8), max(1000, rate * 100.0 / (warmup + steps))))
        hits = dt = 0
        for i in range(warmup + steps):
            lo = (i * n) % max(len(guides_u8) - n, 1)
            rate, hits, dt = cpu.step(guides_u8[lo:lo + n])
            if i >= warmup:
                vals.append(rate)
        cb = cpu.describe(float(np.mean(vals)), n, hits, dt)
        v = cb["value"]
        ms = args.guides / v * 1e3
        workload = f"{args.guides} SpCas9 guides, <=4 mm (core 10 / <=2) vs {gdesc}"
        dtype = "u32 2-bit codes (XOR + popcount per indexed window)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": cb["unit"], "n_gpus": args.gpus,
        "steps": len(vals), "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": workload,
                   "note": "every step is a bounded sample of the workload (a slice of the guides against the whole genome); "
                           "ms_per_step is the projected time of the full job at the measured rate"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": cb["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


if __name__ == "__main__":
    sys.exit(main())
