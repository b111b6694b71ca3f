/*
 * guido_b200.h -- C-ABI of the B200-native guido hot path (libguido_b200.so).
 *
 * Plain C, plain pointers and sizes; no torch / C++ types cross this boundary.  The reference
 * (nkran/guido, pure Python) has no FFI of its own: the boundary is the set of pure functions its
 * public API calls on the hot path.  Each entry point names the reference interface it sits under
 * (paths relative to the reference checkout).  INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *  - every function returns GUIDO_OK (0) or a GUIDO_ERR_* code; guido_last_error() gives the text
 *    (thread-local).  GUIDO_ERR_ARG maps to the ValueError the reference raises for bad arguments,
 *    GUIDO_ERR_PAM to the KeyError of guido/locus.py:178 for a letter outside the IUPAC table.
 *  - output buffers are caller-allocated with a capacity; counts are always exact, and
 *    GUIDO_ERR_CAPACITY means "call again with at least *n elements".
 *  - positions are 0-based.  "global" positions index the concatenated genome coordinate of an
 *    image (see guido_image_contig for the per-contig offsets).
 *  - there is NO CPU fallback: every scan/search entry point fails with GUIDO_ERR_CUDA when no
 *    sm_100 device is usable.
 *  - a handle may be used from one thread at a time.
 */
#ifndef GUIDO_B200_H
#define GUIDO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GUIDO_OK 0
#define GUIDO_ERR_ARG 1
#define GUIDO_ERR_IO 2
#define GUIDO_ERR_CUDA 3
#define GUIDO_ERR_CAPACITY 4
#define GUIDO_ERR_PAM 5
#define GUIDO_ERR_NOMEM 6

/* which windows a PAM scan keeps */
#define GUIDO_SCAN_REGEX 0  /* every PAM match, like re.finditer at guido/locus.py:182-183          */
#define GUIDO_SCAN_GUIDE 1  /* + the 20+P window lies inside its contig and survives the reference's
                               'N' filter (locus.py:208-210,316): '+' drops literal N only, '-' drops
                               any non-ACGT letter (helpers.py:24-26 turns them into N)            */
#define GUIDO_SCAN_STRICT 2 /* + window free of ANY non-ACGT letter on both strands (bowtie's rule
                               for reference positions, used for off-target sites)                  */

/* off-target search algorithm */
#define GUIDO_OT_AUTO 0
#define GUIDO_OT_SCAN 1  /* brute-force: every PAM-valid window against every guide (XOR + popc)     */
#define GUIDO_OT_INDEX 2 /* guide-side seed index: windows look up their core k-mer                  */
#define GUIDO_OT_TABLE 3 /* seed index joined with the genome's PAM-site table (guido_site_table_build) */

typedef struct guido_image guido_image;

typedef struct {
    int64_t n_contigs;
    int64_t total_bases;  /* sum of contig lengths                                   */
    int64_t padded_bases; /* extent of the global coordinate (contigs + pad runs)    */
    int64_t n_runs;       /* non-ACGT runs incl. the pad runs between contigs        */
    int64_t shard_lo;     /* global range whose windows this handle owns on device   */
    int64_t shard_hi;
    int32_t device;       /* -1 while host-only                                      */
    int32_t reserved;
} guido_image_info;

/* one off-target hit, 16 bytes.  hi/lo hold the genomic window in GUIDE orientation as two bit
 * planes (base code A=0,C=1,G=2,T=3; lo = code&1, hi = code>>1): bit i <-> guide position i+1
 * (bits 0..19 protospacer, bits 20.. the genomic PAM).  Bit 31 of `hi` is the guido strand
 * (0 '+', 1 '-').  gpos is the global 0-based position of the window's leftmost base on the
 * forward strand (what bowtie prints as offset, plus the contig's global offset). */
typedef struct {
    uint32_t guide;
    uint32_t gpos;
    uint32_t hi;
    uint32_t lo;
} guido_ot_hit;

typedef struct {
    double kernel_ms;   /* device time of the dominant kernel(s), CUDA events           */
    double total_ms;    /* device time of the whole call incl. copies                   */
    int64_t n_sites;    /* windows examined (PAM-valid, both strands)                   */
    int64_t n_pairs;    /* (window, guide) pairs compared or index entries visited      */
    int64_t n_launches; /* kernels launched by this call                                */
    int64_t bytes_h2d;
    int64_t bytes_d2h;
} guido_stats;

const char *guido_last_error(void);
int guido_version(void);
/* number of usable sm_100 devices (0 when there is none) */
int guido_device_count(void);

/* page-locked host memory for the caller's input/output buffers: device<->host copies into
 * such buffers run at full PCIe rate (any host pointer is accepted everywhere, pageable is just
 * slower).  Pinning is process-wide, so no device needs to be selected. */
int guido_host_alloc(int64_t bytes, void **out);
int guido_host_free(void *p);
/* Peer buffers (multi-GPU merge over NVLink without a collective): device memory that the other
 * ranks of the node can write into.  guido_peer_alloc returns the pointer and a 64-byte handle
 * (cudaIpcMemHandle_t) to pass to the other processes; guido_peer_open maps a peer's buffer into
 * this process (peer access is enabled on first use); guido_copy_dev is an asynchronous
 * device-to-device copy on `stream` -- between GPUs it runs on the copy engines, next to kernels.
 * Under the merge of the per-shard bowtie outputs a multi-GPU run_bowtie needs. */
int guido_peer_alloc(int32_t device, int64_t bytes, void **d_ptr, void *handle64);
int guido_peer_open(int32_t device, const void *handle64, void **d_ptr);
int guido_peer_close(void *d_ptr);
int guido_peer_free(void *d_ptr);
int guido_copy_dev(void *dst, const void *src, int64_t bytes, void *stream);
/* One call for the n copies of a peer exchange: `bytes` from src to dst_ptrs[r] + at_bytes on streams[r]
 * (cudaStream_t each), r = first, first + 1, ... round the ring, each behind after_event (a
 * cudaEvent_t; may be NULL).  guido_streams_join records events[k] on streams[k] and makes
 * onto_stream wait for all of them.  (A step of a few hundred microseconds cannot afford one
 * Python-level call per copy.) */
int guido_peer_send(void *const *dst_ptrs, void *const *streams, int32_t n, int32_t first, int64_t at_bytes,
                    const void *src, int64_t bytes, void *after_event);
int guido_streams_join(void *const *streams, void *const *events, int32_t n, void *onto_stream);

/* device time (ms, CUDA events on the launching stream) of the DOMINANT kernel of the last calls on
 * this handle -- PAM count+fill, or the off-target search kernel -- oldest first, up to 64 kept.
 * The caller must have synchronised the streams it launched on. */
int guido_kernel_times(guido_image *img, double *out_ms, int32_t max_n, int32_t *n);

/* ---- genome image: replaces `bowtie-build` in Genome.build (guido/genome.py:156-171) --------- */
/* FASTA -> two bit planes (1 bit/base each, interleaved per 64 bases) + sorted list of non-ACGT
 * runs + contig table.  Sequence names are the header up to the first whitespace.            */
int guido_image_from_fasta(const char *fasta_path, guido_image **out);
int guido_image_from_seqs(int32_t n, const char *const *names, const char *const *seqs,
                          const int64_t *lens, guido_image **out);
/* seeded synthetic genome (uniform ACGT; n_frac of each contig covered by N runs of
 * run_min..run_max bases).  Bench / test utility; same bytes on every platform. */
int guido_image_synth(int32_t n, const char *const *names, const int64_t *lens, uint64_t seed,
                      double n_frac, int64_t run_min, int64_t run_max, guido_image **out);
int guido_image_save(const guido_image *img, const char *path);
int guido_image_load(const char *path, guido_image **out);
int guido_image_free(guido_image *img);
/* copy shard `shard_ix` of `n_shards` (contiguous, equal block counts, +1 block halo each side)
 * to HBM of `device`; afterwards scans/searches on this handle cover that shard only. */
int guido_image_upload(guido_image *img, int32_t device, int32_t shard_ix, int32_t n_shards);
/* the partition rule itself: global range [lo, hi) of PAM starts owned by shard_ix of n_shards
 * (equal, 65536-base aligned chunks of the padded coordinate; trailing shards may be empty) */
int guido_shard_bounds(int64_t padded_bases, int32_t shard_ix, int32_t n_shards, int64_t *lo, int64_t *hi);
int guido_image_get_info(const guido_image *img, guido_image_info *info);
int guido_image_contig(const guido_image *img, int64_t ix, char *name, int64_t name_cap,
                       int64_t *len, int64_t *global_offset);
/* decode bases [start0, start0+n) of contig ix to upper-case letters (host copy of the image;
 * non-ACGT letters come back as stored in the run list).  Under Genome.sequence / the FASTA
 * fetch of guido/locus.py:773,853. */
int guido_image_fetch(const guido_image *img, int64_t contig_ix, int64_t start0, int64_t n,
                      char *out);

/* the image's non-ACGT runs (N, IUPAC letters, the pads between contigs), sorted: [start, end) in
 * global coordinates, kind 0 = N or pad.  GUIDO_ERR_CAPACITY with *n_runs set when cap is too small.
 * (What a caller needs to tell N-free windows apart without fetching letters; the reference tests
 * `"N" in sequence` on strings, guido/guides.py:312-330.) */
int guido_image_get_runs(const guido_image *img, uint32_t *starts, uint32_t *ends, uint8_t *kinds, int64_t cap,
                         int64_t *n_runs);

/* decode n_windows windows of `width` letters starting at the given GLOBAL positions (host copy);
 * letters outside any contig come back as 'N'.  out: n_windows * width bytes. */
int guido_image_fetch_windows(const guido_image *img, const int64_t *gpos, int64_t n_windows,
                              int32_t width, char *out);

/* ---- PAM scan: under Locus._find_guides_in_interval (guido/locus.py:158-183) ------------------ */
/* host string in, match starts out (ascending; relative to seq[0]); `device` is the CUDA device. */
int guido_pam_scan_seq(const char *seq, int64_t n, const char *pam, int32_t device, int32_t mode,
                       int32_t *fw, int64_t cap_fw, int64_t *n_fw,
                       int32_t *rv, int64_t cap_rv, int64_t *n_rv, guido_stats *stats);
/* same over bases [start0,end0) of a contig of an uploaded image, no sequence upload */
int guido_pam_scan_region(guido_image *img, int64_t contig_ix, int64_t start0, int64_t end0,
                          const char *pam, int32_t mode,
                          int32_t *fw, int64_t cap_fw, int64_t *n_fw,
                          int32_t *rv, int64_t cap_rv, int64_t *n_rv, guido_stats *stats);
/* whole shard, global u32 positions into host buffers */
int guido_pam_scan_genome(guido_image *img, const char *pam, int32_t mode,
                          uint32_t *fw, int64_t cap_fw, int64_t *n_fw,
                          uint32_t *rv, int64_t cap_rv, int64_t *n_rv, guido_stats *stats);
/* whole shard, results stay in HBM: d_fw/d_rv are device pointers (u32 each), d_counts a device
 * pointer to 2 x u64 {n_fw, n_rv}; `stream` is a cudaStream_t (NULL = default).  Asynchronous. */
int guido_pam_scan_genome_dev(guido_image *img, const char *pam, int32_t mode,
                              void *d_fw, int64_t cap_fw, void *d_rv, int64_t cap_rv,
                              void *d_counts, void *stream);

/* ---- off-target search: under run_bowtie (guido/off_targets.py:89-227) ------------------------ */
/* protos: n_guides * 20 letters (the 20-nt protospacers, guide orientation, A/C/G/T).
 * Reports every window w (both strands, inside one contig, free of non-ACGT) whose PAM matches
 * the fixed letters of `pam` exactly and whose protospacer has <= core_mm mismatches in the
 * core_len PAM-proximal bases and <= total_mm + 1 - (#N-like letters of pam) in total -- i.e. the
 * alignments bowtie -n/-l/-e reports that survive guido's PAM filter (off_targets.py:169-198).
 * raw_flags (n_guides bytes, may be NULL) receives 1 for every guide with at least one RAW bowtie
 * alignment (PAM mismatches allowed inside the seed budget): the reference creates the guide's
 * dict key from those (off_targets.py:183-184).  Hits come back sorted by (guide, gpos, strand). */
int guido_offtarget_search(guido_image *img, const char *protos, int64_t n_guides, const char *pam,
                           int32_t core_len, int32_t core_mm, int32_t total_mm, int32_t algo,
                           guido_ot_hit *hits, int64_t cap, int64_t *n_hits, uint8_t *raw_flags,
                           guido_stats *stats);
/* Same search, 8 bytes per hit on the wire instead of 16: hits[i] = guide << 33 | gpos << 1 | strand
 * (the canonical sort key; n_guides < 2^31).  The window planes of guido_ot_hit are a function of
 * (gpos, strand) and the image, so the host re-derives them where it needs them
 * (guido_hits_expand) -- halves the device-to-host copy, which is most of a large search's
 * end-to-end time.  Everything else (order, raw_flags, capacity convention) as above. */
int guido_offtarget_search_packed(guido_image *img, const char *protos, int64_t n_guides, const char *pam,
                                  int32_t core_len, int32_t core_mm, int32_t total_mm, int32_t algo,
                                  uint64_t *hits, int64_t cap, int64_t *n_hits, uint8_t *raw_flags,
                                  guido_stats *stats);

/* The same search with the hit list as CSR by guide, 4.1 bytes per hit on the wire: gpos[i] (cap
 * elements) = global window start of hit i, bit i of strand_bits (cap / 32 + 1 words) = its strand,
 * hits guide_offsets[g] .. guide_offsets[g + 1] belong to guide g (n_guides + 1 entries), inside a
 * guide ordered by (position, strand) -- the order of the other two formats.  What
 * run_bowtie's per-guide dict (guido/off_targets.py:201-225) is built from; guido_hits_expand turns
 * guide << 33 | gpos << 1 | strand back into full records. */
int guido_offtarget_search_csr(guido_image *img, const char *protos, int64_t n_guides, const char *pam,
                               int32_t core_len, int32_t core_mm, int32_t total_mm, int32_t algo,
                               uint32_t *gpos, uint32_t *strand_bits, int64_t cap, int64_t *guide_offsets,
                               int64_t *n_hits, uint8_t *raw_flags, guido_stats *stats);
/* packed records -> guido_ot_hit records, planes cut from the HOST copy of the image exactly as the
 * kernels cut them from HBM (host-only; pam_len = length of the PAM the search used).  Feeds the
 * record building of run_bowtie (guido/off_targets.py:201-225). */
int guido_hits_expand(const guido_image *img, const uint64_t *packed, int64_t n_hits, int32_t pam_len,
                      guido_ot_hit *out);
/* key presence alone: flags[g] = 1 when guide g has at least one RAW bowtie alignment in this
 * handle's stretch of the genome (the relaxed all-window pass guido_offtarget_search runs for the
 * guides without a valid hit; brute force, meant for the few guides that have none anywhere).
 * Ignores a key-range partition of the handle. */
int guido_offtarget_raw_flags(guido_image *img, const char *protos, int64_t n_guides, const char *pam,
                              int32_t core_len, int32_t core_mm, int32_t total_mm, uint8_t *flags);
/* device-resident variant: d_guides = n_guides x {u32 hi, u32 lo} plane words (guide position
 * order, see guido_ot_hit), d_hits = cap x guido_ot_hit, d_count = 1 x u64; unsorted; async. */
int guido_offtarget_search_dev(guido_image *img, const void *d_guides, int64_t n_guides,
                               const char *pam, int32_t core_len, int32_t core_mm, int32_t total_mm,
                               int32_t algo, void *d_hits, int64_t cap, void *d_count, void *stream);
/* device-side guido_encode_guides: d_letters = n_guides x 20 letters in HBM -> d_out = n_guides x
 * {u32 hi, u32 lo}.  d_bad (device u64, may be NULL; set it to ~0 first) receives the smallest
 * guide * 32 + position that holds a letter outside A/C/G/T.  Asynchronous on `stream`. */
int guido_encode_guides_dev(guido_image *img, const void *d_letters, int64_t n_guides, void *d_out, void *d_bad,
                            void *stream);
/* canonical order of a device-resident hit list (e.g. the NCCL-gathered lists of several GPUs):
 * d_hits[0..n_hits) sorted in place by (guide, gpos, strand) -- the order guido_offtarget_search
 * returns and the reference gets from sorting bowtie's rows (off_targets.py:174-225 walks them in
 * file order; guido_b200 fixes the order so 1/2/4/8-GPU outputs are byte-identical).  n_guides
 * bounds the guide field.  Asynchronous on `stream`; scratch memory belongs to the handle. */
int guido_hits_sort_dev(guido_image *img, void *d_hits, int64_t n_hits, int64_t n_guides, void *stream);
/* Merge of per-GPU hit lists after ONE padded NCCL all-gather (no host round trip between the
 * search and the merge): every rank sends `stride` records, record 0 a header whose first 8 bytes
 * hold its hit count (copy d_count there on the device), the hits behind it.  d_gathered =
 * n_parts x stride records; the parts are copied back to back into d_out (cap records) and
 * *d_total (device u64) receives their sum.  A part whose count exceeds stride - 1 is cut (the
 * caller compares the headers with stride afterwards and repeats the step with a larger one).
 * Under the merge of the per-shard bowtie outputs a multi-GPU run_bowtie needs. */
int guido_hits_concat_dev(guido_image *img, const void *d_gathered, int32_t n_parts, int64_t stride,
                          void *d_out, int64_t cap, void *d_total, void *stream);
/* guido_offtarget_search_dev with the handle's key range cut into n_slices (1, 2, 4, 8) equal slices
 * that are indexed and joined one after the other: slice s writes its hits to d_hits + s *
 * cap_per_slice records and its count to d_counts[s] (device u64 x n_slices), and events[s] (a
 * cudaEvent_t each; may be NULL) is recorded on `stream` behind its join -- a multi-GPU caller starts
 * the all-gather of slice s on a second stream while the later slices are still being searched.
 * d_send (may be NULL): behind its join, slice s is also packed into the 8-byte wire format at d_send +
 * s * send_stride words (guido_hits_pack_send_dev: header word = count) BEFORE events[s] is recorded, so
 * that only copies are left for the second stream (kernels would queue behind the join, which fills
 * every SM).  The union of the slices is the hit list of guido_offtarget_search_dev (GUIDO_OT_TABLE only). */
int guido_offtarget_search_sliced_dev(guido_image *img, const void *d_guides, int64_t n_guides, const char *pam,
                                      int32_t core_len, int32_t core_mm, int32_t total_mm, int32_t n_slices,
                                      void *d_hits, int64_t cap_per_slice, void *d_counts, void *const *events,
                                      void *d_send, int64_t send_stride, void *stream);
/* The same merge on the 8-byte wire format (half the bytes over NVLink): guido_hits_pack_send_dev
 * writes *d_count (device u64) into d_send[0] and min(*d_count, cap) packed records behind it;
 * after the all-gather of the `stride`-word slots guido_packed_concat_dev copies the parts back to
 * back (cut parts: as above).  Everything is ordered on `stream`; no host read in between. */
int guido_hits_pack_send_dev(guido_image *img, const void *d_hits, const void *d_count, int64_t cap, void *d_send,
                             void *stream);
int guido_packed_concat_dev(guido_image *img, const void *d_gathered, int32_t n_parts, int64_t stride, void *d_out,
                            int64_t cap, void *d_total, void *stream);
/* device-resident guido_ot_hit records -> the 8-byte wire format of guido_offtarget_search_packed */
int guido_hits_pack_dev(guido_image *img, const void *d_hits, int64_t n_hits, void *d_out, void *stream);
/* Genome-side index of the off-target search, the counterpart of the `bowtie-build` index files
 * (guido/genome.py:156-171): every PAM-valid window of the uploaded shard (both strands, guide
 * orientation) grouped by its core_len PAM-proximal bases, 20 bytes per window in HBM.  Guide
 * independent; depends on the PAM's base sets and on core_len only.  GUIDO_OT_AUTO/TABLE searches
 * build it on first use; this call builds it ahead of time (Genome.build / benchmarks) and reports
 * its size and build time.  GUIDO_ERR_NOMEM: more than 2^31 windows or not enough free HBM
 * (AUTO then streams the genome through the seed index instead). */
int guido_site_table_build(guido_image *img, const char *pam, int32_t core_len, int64_t *n_sites,
                           double *build_ms);
int guido_site_table_free(guido_image *img);
/* Multi-GPU off-target search without replicated work: instead of giving every process a stretch of
 * the genome (guido_image_upload shards; every process then needs the whole guide index), give
 * every process the WHOLE image and 1/n_parts of the core-key space.  The handle then tables only
 * the windows whose core k-mer falls into part `part_ix` (top log2(n_parts) bits of the key;
 * n_parts a power of two <= 32), builds only that part of the guide index per search, and its
 * TABLE/AUTO searches return exactly the hits of those windows: the union over the parts is the
 * hit list of the image.  Other algorithms refuse to run on a partitioned handle, and raw_flags of
 * guido_offtarget_search then only mean "has a valid hit in this part" (OR them over the parts and
 * ask an unpartitioned handle about the guides that are left). */
int guido_site_table_partition(guido_image *img, int32_t part_ix, int32_t n_parts);
/* encode n x 20 letters into the {hi, lo} plane words the _dev call takes (host helper) */
int guido_encode_guides(const char *protos, int64_t n_guides, uint32_t *out_hi_lo);

/* ---- MMEJ: under generate_mmej_patterns (guido/mmej.py:8-64) ---------------------------------- */
/* windows: concatenated upper-case letters, offsets[n+1]; cuts[n] relative cut positions.
 * out_tuples receives the kept candidates per window, in the reference's order, packed
 * left_start | left_end<<8 | right_start<<16 | right_end<<24; out_offsets[n+1] delimits them.
 * n_cand[n] (may be NULL) = candidates before nested-pattern elimination (0 => the reference
 * returns None).  Each side of the cut may be at most 255 letters. */
int guido_mmej_batch(const char *windows, const int64_t *offsets, const int32_t *cuts, int64_t n,
                     int32_t device, uint32_t *out_tuples, int64_t cap, int64_t *out_offsets,
                     int32_t *n_cand, guido_stats *stats);

/* The same enumeration for windows of the resident image: window g = global positions
 * [starts[g], starts[g] + lens[g]) (forward strand, the orientation of Guide.locus_seq,
 * guido/guides.py:69-84), cuts[g] relative to its start.  Nothing but 12 bytes per site goes up;
 * windows that touch a non-ACGT run are read from the host copy of the image and take the general
 * kernel.  Replaces the per-guide sequence[slice] + generate_mmej_patterns of
 * guido/guides.py:132-182 when the guides come from a built Genome. */
int guido_mmej_sites(const guido_image *img, const uint32_t *starts, const int32_t *lens, const int32_t *cuts,
                     int64_t n, uint32_t *out_tuples, int64_t cap, int64_t *out_offsets, int32_t *n_cand,
                     guido_stats *stats);

#ifdef __cplusplus
}
#endif
#endif /* GUIDO_B200_H */
