"""ctypes binding of libguido_b200.so (the C-ABI declared in include/guido_b200.h).

There is no CPU fallback: if the shared library is missing, or no sm_100 device is usable,
every accelerated call raises.  Arrays cross the boundary as numpy buffers (plain pointers).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# GUIDO_B200_LIB: another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("GUIDO_B200_LIB") or os.path.join(_HERE, "lib", "libguido_b200.so")

OK, ERR_ARG, ERR_IO, ERR_CUDA, ERR_CAPACITY, ERR_PAM, ERR_NOMEM = range(7)
SCAN_REGEX, SCAN_GUIDE, SCAN_STRICT = 0, 1, 2
OT_AUTO, OT_SCAN, OT_INDEX, OT_TABLE = 0, 1, 2, 3

OT_HIT_DTYPE = np.dtype([("guide", "<u4"), ("gpos", "<u4"), ("hi", "<u4"), ("lo", "<u4")])


class ImageInfo(ctypes.Structure):
    _fields_ = [("n_contigs", ctypes.c_int64), ("total_bases", ctypes.c_int64),
                ("padded_bases", ctypes.c_int64), ("n_runs", ctypes.c_int64),
                ("shard_lo", ctypes.c_int64), ("shard_hi", ctypes.c_int64),
                ("device", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class Stats(ctypes.Structure):
    _fields_ = [("kernel_ms", ctypes.c_double), ("total_ms", ctypes.c_double),
                ("n_sites", ctypes.c_int64), ("n_pairs", ctypes.c_int64),
                ("n_launches", ctypes.c_int64), ("bytes_h2d", ctypes.c_int64),
                ("bytes_d2h", ctypes.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None
_P = ctypes.POINTER
_vp, _i64, _i32, _cp = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_char_p

_SIGNATURES = {
    "guido_last_error": (ctypes.c_char_p, []),
    "guido_version": (ctypes.c_int, []),
    "guido_device_count": (ctypes.c_int, []),
    "guido_host_alloc": (ctypes.c_int, [_i64, _P(_vp)]),
    "guido_host_free": (ctypes.c_int, [_vp]),
    "guido_peer_alloc": (ctypes.c_int, [_i32, _i64, _P(_vp), _vp]),
    "guido_peer_open": (ctypes.c_int, [_i32, _vp, _P(_vp)]),
    "guido_peer_close": (ctypes.c_int, [_vp]),
    "guido_peer_free": (ctypes.c_int, [_vp]),
    "guido_copy_dev": (ctypes.c_int, [_vp, _vp, _i64, _vp]),
    "guido_peer_send": (ctypes.c_int, [_vp, _vp, _i32, _i32, _i64, _vp, _i64, _vp]),
    "guido_streams_join": (ctypes.c_int, [_vp, _vp, _i32, _vp]),
    "guido_kernel_times": (ctypes.c_int, [_vp, _vp, _i32, _P(_i32)]),
    "guido_image_from_fasta": (ctypes.c_int, [_cp, _P(_vp)]),
    "guido_image_from_seqs": (ctypes.c_int, [_i32, _P(_cp), _P(_cp), _P(_i64), _P(_vp)]),
    "guido_image_synth": (ctypes.c_int, [_i32, _P(_cp), _P(_i64), ctypes.c_uint64, ctypes.c_double,
                                         _i64, _i64, _P(_vp)]),
    "guido_image_save": (ctypes.c_int, [_vp, _cp]),
    "guido_image_load": (ctypes.c_int, [_cp, _P(_vp)]),
    "guido_image_free": (ctypes.c_int, [_vp]),
    "guido_image_upload": (ctypes.c_int, [_vp, _i32, _i32, _i32]),
    "guido_shard_bounds": (ctypes.c_int, [_i64, _i32, _i32, _P(_i64), _P(_i64)]),
    "guido_image_get_info": (ctypes.c_int, [_vp, _P(ImageInfo)]),
    "guido_image_contig": (ctypes.c_int, [_vp, _i64, _cp, _i64, _P(_i64), _P(_i64)]),
    "guido_image_fetch": (ctypes.c_int, [_vp, _i64, _i64, _i64, _vp]),
    "guido_image_fetch_windows": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp]),
    "guido_pam_scan_seq": (ctypes.c_int, [_cp, _i64, _cp, _i32, _i32, _vp, _i64, _P(_i64), _vp, _i64,
                                          _P(_i64), _P(Stats)]),
    "guido_pam_scan_region": (ctypes.c_int, [_vp, _i64, _i64, _i64, _cp, _i32, _vp, _i64, _P(_i64),
                                             _vp, _i64, _P(_i64), _P(Stats)]),
    "guido_pam_scan_genome": (ctypes.c_int, [_vp, _cp, _i32, _vp, _i64, _P(_i64), _vp, _i64,
                                             _P(_i64), _P(Stats)]),
    "guido_pam_scan_genome_dev": (ctypes.c_int, [_vp, _cp, _i32, _vp, _i64, _vp, _i64, _vp, _vp]),
    "guido_offtarget_search": (ctypes.c_int, [_vp, _cp, _i64, _cp, _i32, _i32, _i32, _i32, _vp, _i64,
                                              _P(_i64), _vp, _P(Stats)]),
    "guido_offtarget_search_packed": (ctypes.c_int, [_vp, _cp, _i64, _cp, _i32, _i32, _i32, _i32, _vp, _i64,
                                                     _P(_i64), _vp, _P(Stats)]),
    "guido_offtarget_search_csr": (ctypes.c_int, [_vp, _vp, _i64, _cp, _i32, _i32, _i32, _i32, _vp, _vp, _i64, _vp,
                                                  _P(_i64), _vp, _P(Stats)]),
    "guido_hits_expand": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp]),
    "guido_offtarget_search_dev": (ctypes.c_int, [_vp, _vp, _i64, _cp, _i32, _i32, _i32, _i32, _vp,
                                                  _i64, _vp, _vp]),
    "guido_offtarget_raw_flags": (ctypes.c_int, [_vp, _cp, _i64, _cp, _i32, _i32, _i32, _vp]),
    "guido_hits_sort_dev": (ctypes.c_int, [_vp, _vp, _i64, _i64, _vp]),
    "guido_hits_concat_dev": (ctypes.c_int, [_vp, _vp, _i32, _i64, _vp, _i64, _vp, _vp]),
    "guido_hits_pack_dev": (ctypes.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "guido_offtarget_search_sliced_dev": (ctypes.c_int, [_vp, _vp, _i64, _cp, _i32, _i32, _i32, _i32, _vp, _i64, _vp, _vp, _vp,
                                                         _i64, _vp]),
    "guido_hits_pack_send_dev": (ctypes.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "guido_packed_concat_dev": (ctypes.c_int, [_vp, _vp, _i32, _i64, _vp, _i64, _vp, _vp]),
    "guido_encode_guides_dev": (ctypes.c_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "guido_site_table_build": (ctypes.c_int, [_vp, _cp, _i32, _P(_i64), _P(ctypes.c_double)]),
    "guido_site_table_free": (ctypes.c_int, [_vp]),
    "guido_site_table_partition": (ctypes.c_int, [_vp, _i32, _i32]),
    "guido_encode_guides": (ctypes.c_int, [_cp, _i64, _vp]),
    "guido_mmej_batch": (ctypes.c_int, [_cp, _vp, _vp, _i64, _i32, _vp, _i64, _vp, _vp, _P(Stats)]),
    "guido_image_get_runs": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _P(ctypes.c_int64)]),
    "guido_mmej_sites": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _P(Stats)]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """The loaded shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (or `make -C guido_b200/csrc`).  guido_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    return lib().guido_last_error().decode(errors="replace")


def check(rc):
    if rc == OK:
        return
    msg = last_error()
    if rc == ERR_ARG:
        raise ValueError(msg)
    if rc == ERR_PAM:
        raise KeyError(msg.strip("'"))
    if rc == ERR_IO:
        raise OSError(msg)
    if rc == ERR_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def device_count():
    return int(lib().guido_device_count())


def default_device():
    """GUIDO_DEVICE, else LOCAL_RANK (one process per GPU under torchrun), else 0."""
    for key in ("GUIDO_DEVICE", "LOCAL_RANK"):
        if os.environ.get(key, "") != "":
            return int(os.environ[key])
    return 0


def shard_bounds(padded_bases, shard_ix, n_shards):
    """(lo, hi): global PAM-start range owned by shard_ix of n_shards (the library's partition rule)."""
    lo, hi = ctypes.c_int64(), ctypes.c_int64()
    check(lib().guido_shard_bounds(int(padded_bases), shard_ix, n_shards, ctypes.byref(lo), ctypes.byref(hi)))
    return lo.value, hi.value


def _ptr(a):
    return a.ctypes.data if a is not None else None


class PinnedBuffer:
    """Page-locked host memory exposed as numpy arrays (grown on demand, freed with the owner)."""

    def __init__(self):
        self._p = ctypes.c_void_p()
        self._cap = 0

    def array(self, n, dtype):
        dtype = np.dtype(dtype)
        need = max(int(n) * dtype.itemsize, 1)
        if need > self._cap:
            self.free()
            cap = max(need * 5 // 4, 1 << 20)
            check(lib().guido_host_alloc(cap, ctypes.byref(self._p)))
            self._cap = cap
        buf = (ctypes.c_char * need).from_address(self._p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(n))

    def free(self):
        if self._p:
            lib().guido_host_free(self._p)
            self._p = ctypes.c_void_p()
            self._cap = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Image:
    """Owner of one ``guido_image`` handle (host planes + optional HBM copy of one shard)."""

    def __init__(self, handle):
        self._h = ctypes.c_void_p(handle)
        self._contigs = None
        self._pins = {}

    # ---- construction -----------------------------------------------------------------
    @classmethod
    def from_fasta(cls, path):
        h = ctypes.c_void_p()
        check(lib().guido_image_from_fasta(str(path).encode(), ctypes.byref(h)))
        return cls(h.value)

    @classmethod
    def from_seqs(cls, contigs):
        """contigs: iterable of (name, sequence str/bytes)."""
        contigs = list(contigs)
        n = len(contigs)
        names = (ctypes.c_char_p * n)(*[str(c[0]).encode() for c in contigs])
        blobs = [c[1].encode() if isinstance(c[1], str) else bytes(c[1]) for c in contigs]
        seqs = (ctypes.c_char_p * n)(*blobs)
        lens = (ctypes.c_int64 * n)(*[len(b) for b in blobs])
        h = ctypes.c_void_p()
        check(lib().guido_image_from_seqs(n, names, seqs, lens, ctypes.byref(h)))
        return cls(h.value)

    @classmethod
    def synth(cls, lens, seed, n_frac=0.005, run_min=50, run_max=5000, names=None):
        n = len(lens)
        names_arr = (ctypes.c_char_p * n)(*[s.encode() for s in names]) if names else None
        lens_arr = (ctypes.c_int64 * n)(*[int(x) for x in lens])
        h = ctypes.c_void_p()
        check(lib().guido_image_synth(n, names_arr, lens_arr, int(seed), float(n_frac), int(run_min),
                                      int(run_max), ctypes.byref(h)))
        return cls(h.value)

    @classmethod
    def load(cls, path):
        h = ctypes.c_void_p()
        check(lib().guido_image_load(str(path).encode(), ctypes.byref(h)))
        return cls(h.value)

    def save(self, path):
        check(lib().guido_image_save(self._h, str(path).encode()))

    def close(self):
        for pin in getattr(self, "_pins", {}).values():
            pin.free()
        if self._h:
            lib().guido_image_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- metadata ---------------------------------------------------------------------
    def upload(self, device=None, shard_ix=0, n_shards=1):
        check(lib().guido_image_upload(self._h, default_device() if device is None else device,
                                       shard_ix, n_shards))
        return self

    @property
    def info(self):
        inf = ImageInfo()
        check(lib().guido_image_get_info(self._h, ctypes.byref(inf)))
        return inf

    @property
    def contigs(self):
        """[(name, length, global_offset), ...]"""
        if self._contigs is None:
            out = []
            buf = ctypes.create_string_buffer(4096)
            ln, off = ctypes.c_int64(), ctypes.c_int64()
            for i in range(self.info.n_contigs):
                check(lib().guido_image_contig(self._h, i, buf, 4096, ctypes.byref(ln), ctypes.byref(off)))
                out.append((buf.value.decode(), ln.value, off.value))
            self._contigs = out
        return self._contigs

    def contig_index(self, name):
        for i, c in enumerate(self.contigs):
            if c[0] == name:
                return i
        raise KeyError(name)

    def locate(self, gpos):
        """global positions -> (contig index array, offset-in-contig array)."""
        gpos = np.asarray(gpos, dtype=np.int64)
        offs = np.array([c[2] for c in self.contigs], dtype=np.int64)
        ix = np.searchsorted(offs, gpos, side="right") - 1
        return ix, gpos - offs[ix]

    def fetch(self, contig_ix, start0, n):
        buf = np.empty(n, dtype=np.uint8)
        check(lib().guido_image_fetch(self._h, contig_ix, start0, n, buf.ctypes.data))
        return buf.tobytes().decode()

    def kernel_times(self, max_n=64):
        """ms of the dominant kernel of the last calls (oldest first); synchronise first."""
        out = np.zeros(max_n, np.float64)
        n = ctypes.c_int32()
        check(lib().guido_kernel_times(self._h, out.ctypes.data, max_n, ctypes.byref(n)))
        return out[:n.value]

    def fetch_windows(self, gpos, width):
        """(n, width) uint8 array of letters for windows starting at global positions gpos."""
        gpos = np.ascontiguousarray(gpos, dtype=np.int64)
        out = np.empty((len(gpos), width), np.uint8)
        check(lib().guido_image_fetch_windows(self._h, gpos.ctypes.data, len(gpos), width, out.ctypes.data))
        return out

    def runs(self):
        """(starts, ends, kinds) of the non-ACGT runs (incl. the pads between contigs), sorted by start"""
        n = ctypes.c_int64()
        cap = max(int(self.info.n_runs), 1)
        st, en, kd = np.empty(cap, np.uint32), np.empty(cap, np.uint32), np.empty(cap, np.uint8)
        check(lib().guido_image_get_runs(self._h, st.ctypes.data, en.ctypes.data, kd.ctypes.data, cap, ctypes.byref(n)))
        return st[:n.value], en[:n.value], kd[:n.value]

    def mmej_sites(self, starts, lens, cuts, stats=None, cap_hint=None, pinned=False):
        """MMEJ candidates of windows of the RESIDENT image: window g = global positions
        [starts[g], starts[g] + lens[g]) (forward strand), cuts[g] relative to its start.
        Returns (tuples, offsets, n_cand) like :func:`mmej_batch`; only 12 bytes per site go up."""
        starts = np.ascontiguousarray(starts, dtype=np.uint32)
        n = len(starts)
        lens = np.ascontiguousarray(np.broadcast_to(np.asarray(lens, dtype=np.int32), (n,)))
        cuts = np.ascontiguousarray(np.broadcast_to(np.asarray(cuts, dtype=np.int32), (n,)))
        out_off = np.zeros(n + 1, np.int64)
        n_cand = np.zeros(n, np.int32)
        st = stats if stats is not None else Stats()
        cap = int(cap_hint) if cap_hint else max(1024, 320 * n)
        while True:
            # pinned=True: a reused page-locked buffer (valid until the next call), like pam_scan_genome's
            out = self._pinned("mmej", cap, np.uint32) if pinned else np.empty(cap, np.uint32)
            rc = lib().guido_mmej_sites(self._h, starts.ctypes.data, lens.ctypes.data, cuts.ctypes.data, n, out.ctypes.data,
                                        cap, out_off.ctypes.data, n_cand.ctypes.data, ctypes.byref(st))
            if rc == ERR_CAPACITY:
                cap = int(out_off[n])
                continue
            check(rc)
            return out[:out_off[n]], out_off, n_cand

    # ---- PAM scan ---------------------------------------------------------------------
    def _pinned(self, name, n, dtype):
        return self._pins.setdefault(name, PinnedBuffer()).array(n, dtype)

    def pam_scan_genome(self, pam, mode=SCAN_GUIDE, stats=None, cap_hint=None, pinned=False):
        """Whole shard -> (fw, rv) uint32 arrays of global PAM-start positions (ascending)."""
        span = self.info.shard_hi - self.info.shard_lo
        cap = int(cap_hint) if cap_hint else max(1024, span // 8)
        st = stats if stats is not None else Stats()
        while True:
            if pinned:  # views of page-locked buffers owned by this image: valid until the next call
                fw, rv = self._pinned("fw", cap, np.uint32), self._pinned("rv", cap, np.uint32)
            else:
                fw, rv = np.empty(cap, np.uint32), np.empty(cap, np.uint32)
            nf, nr = ctypes.c_int64(), ctypes.c_int64()
            rc = lib().guido_pam_scan_genome(self._h, pam.encode(), mode, fw.ctypes.data, cap,
                                             ctypes.byref(nf), rv.ctypes.data, cap, ctypes.byref(nr),
                                             ctypes.byref(st))
            if rc == ERR_CAPACITY:
                cap = max(nf.value, nr.value)
                continue
            check(rc)
            return fw[:nf.value], rv[:nr.value]

    def pam_scan_region(self, contig_ix, start0, end0, pam, mode=SCAN_REGEX, stats=None):
        n = max(end0 - start0, 0)
        cap = max(256, n // 6)
        st = stats if stats is not None else Stats()
        while True:
            fw, rv = np.empty(cap, np.int32), np.empty(cap, np.int32)
            nf, nr = ctypes.c_int64(), ctypes.c_int64()
            rc = lib().guido_pam_scan_region(self._h, contig_ix, start0, end0, pam.encode(), mode,
                                             fw.ctypes.data, cap, ctypes.byref(nf), rv.ctypes.data, cap,
                                             ctypes.byref(nr), ctypes.byref(st))
            if rc == ERR_CAPACITY:
                cap = max(nf.value, nr.value)
                continue
            check(rc)
            return fw[:nf.value], rv[:nr.value]

    def pam_scan_genome_dev(self, pam, mode, d_fw, cap_fw, d_rv, cap_rv, d_counts, stream=0):
        check(lib().guido_pam_scan_genome_dev(self._h, pam.encode(), mode, d_fw, cap_fw, d_rv, cap_rv,
                                              d_counts, stream))

    # ---- off-target search ------------------------------------------------------------
    def site_table_build(self, pam="NGG", core_length=10):
        """Build (or reuse) the genome-side PAM-site table for this PAM / core length -- the
        counterpart of bowtie-build's index files.  Returns (n_sites, build_ms; 0.0 if reused)."""
        n, ms = ctypes.c_int64(), ctypes.c_double()
        check(lib().guido_site_table_build(self._h, pam.encode(), core_length, ctypes.byref(n), ctypes.byref(ms)))
        return n.value, ms.value

    def site_table_free(self):
        check(lib().guido_site_table_free(self._h))

    def site_table_partition(self, part_ix, n_parts):
        """Restrict this handle's PAM-site table -- and therefore its TABLE/AUTO searches -- to part
        ``part_ix`` of ``n_parts`` (a power of two) of the core-key space; see the header."""
        check(lib().guido_site_table_partition(self._h, part_ix, n_parts))
        return self

    def offtarget_search(self, protos, pam="NGG", core_length=10, core_mismatches=2,
                         total_mismatches=4, algo=OT_AUTO, want_raw_flags=True, stats=None, pinned=False,
                         packed=False):
        """protos: list of 20-nt strings (or an (n,20) uint8 array of letters).

        Returns (hits structured array sorted by (guide, gpos, strand), raw_flags uint8[n] | None).
        ``packed=True``: hits come back as uint64 ``guide << 33 | gpos << 1 | strand`` (half the
        device-to-host bytes); ``expand_hits`` turns them into the structured records on the host.
        """
        if isinstance(protos, np.ndarray):
            blob = np.ascontiguousarray(protos, dtype=np.uint8).tobytes()
            n = protos.shape[0]
        else:
            protos = list(protos)
            n = len(protos)
            for p in protos:
                if len(p) != 20:
                    raise ValueError(f"protospacer must be 20 nt, got {len(p)}: {p!r}")
            blob = "".join(protos).encode()
        flags = np.zeros(n, np.uint8) if want_raw_flags else None
        st = stats if stats is not None else Stats()
        cap = max(4096, 32 * n, getattr(self, "_ot_hint", 0))
        dtype = np.uint64 if packed else OT_HIT_DTYPE
        call = lib().guido_offtarget_search_packed if packed else lib().guido_offtarget_search
        while True:
            hits = self._pinned("hits", cap, dtype) if pinned else np.empty(cap, dtype)
            nh = ctypes.c_int64()
            rc = call(self._h, blob, n, pam.encode(), core_length, core_mismatches, total_mismatches, algo,
                      hits.ctypes.data, cap, ctypes.byref(nh), _ptr(flags), ctypes.byref(st))
            if rc == ERR_CAPACITY:
                cap = nh.value
                continue
            check(rc)
            self._ot_hint = nh.value + nh.value // 16 + 16  # right-size the next call's buffer
            return hits[:nh.value], flags

    def offtarget_search_csr(self, protos, pam="NGG", core_length=10, core_mismatches=2, total_mismatches=4,
                             algo=OT_AUTO, want_raw_flags=True, stats=None, pinned=False):
        """The search with the hit list as CSR by guide -- 4.1 bytes per hit over PCIe instead of 8 or 16:
        returns (offsets int64[n + 1], gpos uint32[n_hits], strand_bits uint32[ceil(n_hits / 32)], raw_flags).
        Hits offsets[g] .. offsets[g + 1] belong to guide g, ordered by (position, strand);
        ``csr_to_packed`` / ``expand_hits`` rebuild the other formats on the host."""
        if isinstance(protos, np.ndarray):
            keep = np.ascontiguousarray(protos, dtype=np.uint8)   # letters are read in place: no copy of the guide set
            blob, n = keep.ctypes.data, protos.shape[0]
        else:
            protos = list(protos)
            n = len(protos)
            for p in protos:
                if len(p) != 20:
                    raise ValueError(f"protospacer must be 20 nt, got {len(p)}: {p!r}")
            blob = "".join(protos).encode()
        flags = np.zeros(n, np.uint8) if want_raw_flags else None
        st = stats if stats is not None else Stats()
        cap = max(4096, 32 * n, getattr(self, "_ot_hint", 0))
        offsets = np.empty(n + 1, np.int64)
        while True:
            gpos = self._pinned("csr_gpos", cap, np.uint32) if pinned else np.empty(cap, np.uint32)
            bits = self._pinned("csr_bits", cap // 32 + 1, np.uint32) if pinned else np.empty(cap // 32 + 1, np.uint32)
            nh = ctypes.c_int64()
            rc = lib().guido_offtarget_search_csr(self._h, blob, n, pam.encode(), core_length, core_mismatches, total_mismatches,
                                                  algo, gpos.ctypes.data, bits.ctypes.data, cap, offsets.ctypes.data,
                                                  ctypes.byref(nh), _ptr(flags), ctypes.byref(st))
            if rc == ERR_CAPACITY:
                cap = nh.value
                continue
            check(rc)
            self._ot_hint = nh.value + nh.value // 16 + 16
            return offsets, gpos[:nh.value], bits[:(nh.value + 31) // 32], flags

    @staticmethod
    def csr_to_packed(offsets, gpos, strand_bits):
        """CSR hit list -> uint64 ``guide << 33 | gpos << 1 | strand`` (the packed format, same order)."""
        n = len(gpos)
        guide = np.repeat(np.arange(len(offsets) - 1, dtype=np.uint64), np.diff(offsets))
        strand = (np.unpackbits(np.ascontiguousarray(strand_bits).view(np.uint8), bitorder="little")[:n]).astype(np.uint64)
        return (guide << np.uint64(33)) | (gpos.astype(np.uint64) << np.uint64(1)) | strand

    def expand_hits(self, packed, pam_len=3):
        """uint64 packed hits -> structured records (window planes cut from the host image)."""
        packed = np.ascontiguousarray(packed, dtype=np.uint64)
        out = np.empty(len(packed), OT_HIT_DTYPE)
        check(lib().guido_hits_expand(self._h, packed.ctypes.data, len(packed), pam_len, out.ctypes.data))
        return out

    def offtarget_raw_flags(self, protos, pam="NGG", core_length=10, core_mismatches=2, total_mismatches=4):
        """uint8[n]: 1 where the guide has any raw bowtie alignment in this handle's stretch of the
        genome (brute force over every window; for the few guides without a valid hit)."""
        protos = list(protos)
        for p in protos:
            if len(p) != 20:
                raise ValueError(f"protospacer must be 20 nt, got {len(p)}: {p!r}")
        flags = np.zeros(len(protos), np.uint8)
        check(lib().guido_offtarget_raw_flags(self._h, "".join(protos).encode(), len(protos), pam.encode(),
                                              core_length, core_mismatches, total_mismatches, _ptr(flags)))
        return flags

    def offtarget_search_dev(self, d_guides, n_guides, pam, core_length, core_mismatches,
                             total_mismatches, algo, d_hits, cap, d_count, stream=0):
        check(lib().guido_offtarget_search_dev(self._h, d_guides, n_guides, pam.encode(), core_length,
                                               core_mismatches, total_mismatches, algo, d_hits, cap,
                                               d_count, stream))

    def offtarget_search_sliced_dev(self, d_guides, n_guides, pam, core_length, core_mismatches, total_mismatches,
                                    n_slices, d_hits, cap_per_slice, d_counts, events=None, stream=0, d_send=None,
                                    send_stride=0):
        """the table search slice by slice of the key range: hits of slice s at d_hits + s * cap_per_slice records,
        count in d_counts[s]; events: raw cudaEvent_t handles (ints), recorded behind each slice's join (and behind
        its packing into d_send + s * send_stride words when d_send is given)"""
        ev = None
        if events is not None:
            ev = (ctypes.c_void_p * n_slices)(*[ctypes.c_void_p(int(e)) for e in events])
        check(lib().guido_offtarget_search_sliced_dev(self._h, d_guides, n_guides, pam.encode(), core_length, core_mismatches,
                                                      total_mismatches, n_slices, d_hits, cap_per_slice, d_counts, ev, d_send,
                                                      send_stride, stream))

    def encode_guides_dev(self, d_letters, n_guides, d_out, d_bad=None, stream=0):
        """n x 20 letters in HBM -> n x {hi, lo} plane words in HBM; asynchronous"""
        check(lib().guido_encode_guides_dev(self._h, d_letters, n_guides, d_out, d_bad, stream))

    def concat_hits_dev(self, d_gathered, n_parts, stride, d_out, cap, d_total, stream=0):
        """padded, header-prefixed per-rank lists (one all-gather) -> one contiguous list; asynchronous"""
        check(lib().guido_hits_concat_dev(self._h, d_gathered, n_parts, stride, d_out, cap, d_total, stream))

    def pack_send_dev(self, d_hits, d_count, cap, d_send, stream=0):
        """hits -> 8-byte wire records behind a header word that holds *d_count (read on the device); asynchronous"""
        check(lib().guido_hits_pack_send_dev(self._h, d_hits, d_count, cap, d_send, stream))

    def concat_packed_dev(self, d_gathered, n_parts, stride, d_out, cap, d_total, stream=0):
        """concat_hits_dev for slots of 8-byte wire records; asynchronous"""
        check(lib().guido_packed_concat_dev(self._h, d_gathered, n_parts, stride, d_out, cap, d_total, stream))

    def pack_hits_dev(self, d_hits, n_hits, d_out, stream=0):
        """device-resident 16-byte records -> 8-byte guide << 33 | gpos << 1 | strand; asynchronous"""
        check(lib().guido_hits_pack_dev(self._h, d_hits, n_hits, d_out, stream))

    def sort_hits_dev(self, d_hits, n_hits, n_guides, stream=0):
        """device-resident hit records -> (guide, gpos, strand) order, in place, asynchronous"""
        check(lib().guido_hits_sort_dev(self._h, d_hits, n_hits, n_guides, stream))


def pam_scan_seq(seq, pam, device=None, mode=SCAN_REGEX, stats=None):
    """(pams_fw, pams_rv): 0-based match starts in ``seq`` -- what the two re.finditer passes of
    guido/locus.py:182-183 return, computed on the GPU."""
    b = seq.encode() if isinstance(seq, str) else bytes(seq)
    n = len(b)
    cap = max(256, n // 6)
    st = stats if stats is not None else Stats()
    dev = default_device() if device is None else device
    while True:
        fw, rv = np.empty(cap, np.int32), np.empty(cap, np.int32)
        nf, nr = ctypes.c_int64(), ctypes.c_int64()
        rc = lib().guido_pam_scan_seq(b, n, pam.encode(), dev, mode, fw.ctypes.data, cap, ctypes.byref(nf),
                                      rv.ctypes.data, cap, ctypes.byref(nr), ctypes.byref(st))
        if rc == ERR_CAPACITY:
            cap = max(nf.value, nr.value)
            continue
        check(rc)
        return fw[:nf.value], rv[:nr.value]


def encode_guides(protos):
    """list of 20-nt strings (or an (n,20) uint8 array of letters) -> (n,2) uint32 {hi, lo} plane
    words (guide position order)."""
    if isinstance(protos, np.ndarray):
        blob, n = np.ascontiguousarray(protos, dtype=np.uint8).tobytes(), protos.shape[0]
    else:
        protos = list(protos)
        blob, n = "".join(protos).encode(), len(protos)
    out = np.empty((n, 2), np.uint32)
    check(lib().guido_encode_guides(blob, n, out.ctypes.data))
    return out


def mmej_batch(windows, cuts, device=None, stats=None):
    """windows: list of str; cuts: list of int.

    Returns (tuples uint32[total] packed ls|le<<8|rs<<16|re<<24, offsets int64[n+1],
    n_cand int32[n])."""
    n = len(windows)
    offsets = np.zeros(n + 1, np.int64)
    np.cumsum([len(w) for w in windows], out=offsets[1:])
    blob = "".join(windows).encode()
    cuts = np.ascontiguousarray(cuts, dtype=np.int32)
    out_off = np.zeros(n + 1, np.int64)
    n_cand = np.zeros(n, np.int32)
    st = stats if stats is not None else Stats()
    dev = default_device() if device is None else device
    cap = max(1024, 512 * n)
    while True:
        out = np.empty(cap, np.uint32)
        rc = lib().guido_mmej_batch(blob, offsets.ctypes.data, cuts.ctypes.data, n, dev, out.ctypes.data,
                                    cap, out_off.ctypes.data, n_cand.ctypes.data, ctypes.byref(st))
        if rc == ERR_CAPACITY:
            cap = int(out_off[n])
            continue
        check(rc)
        return out[:out_off[n]], out_off, n_cand
