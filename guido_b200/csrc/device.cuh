// device.cuh -- device-side state of an uploaded genome image and helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>

#include "internal.h"

namespace guido {

#define GUIDO_CUDA(call)                                                                   \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_), __FILE__, __LINE__, \
                      cudaGetErrorString(e_));                                             \
            return GUIDO_ERR_CUDA;                                                         \
        }                                                                                  \
    } while (0)

struct DeviceState {
    int device = -1;
    int sm_count = 0;
    int scan_grid[3] = {0, 0, 0};  // resident-CTA grids of the PAM count / fill / one-pass kernels (cached)
    int64_t blk_lo = 0;  // first resident block (global block index), includes the halo
    int64_t blk_n = 0;   // resident blocks (+1 spare at the end so `next` loads stay in bounds)
    ulonglong2 *d_planes = nullptr;
    uint32_t *d_run_start = nullptr, *d_run_end = nullptr;
    uint8_t *d_run_kind = nullptr;
    uint32_t *d_tile_first_run = nullptr;  // per 65536-base tile: first run ending after tile start - 32
    int n_runs = 0;
    // capacities (bytes) of the five arrays above: a handle that is uploaded again -- or the cached
    // per-device context of the sequence-level entry points -- reuses them
    int64_t planes_cap = 0, run_start_cap = 0, run_end_cap = 0, run_kind_cap = 0, tile_first_run_cap = 0;
    int64_t shard_lo = 0, shard_hi = 0;  // owned global positions (PAM starts), 64-aligned
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // ring of event pairs around the dominant kernel of the last calls (guido_kernel_times)
    static constexpr int kKev = 64;
    cudaEvent_t kev[kKev][2] = {};
    long long kev_count = 0;
    // scratch (grown on demand, reused between calls)
    unsigned long long *d_desc = nullptr;  int64_t desc_cap = 0;     // look-back descriptors
    unsigned long long *d_counts = nullptr;                          // 8 x u64 counters
    unsigned int *d_tile_counter = nullptr;
    void *d_out[2] = {nullptr, nullptr};   int64_t out_cap[2] = {0, 0};  // bytes
    void *d_guides = nullptr;              int64_t guides_cap = 0;       // bytes
    void *d_flags = nullptr;               int64_t flags_cap = 0;
    void *d_index = nullptr;               int64_t index_cap = 0;        // OT seed index entries
    void *d_bucket = nullptr;              int64_t bucket_cap = 0;       // OT seed index offsets
    void *h_pinned = nullptr;              int64_t pinned_cap = 0;
    int64_t hits_hint = 0;  // hit count of the previous off-target search (sizes the next buffer)
    // PAM-site table: every PAM-valid window of the resident shard (guide orientation), grouped by
    // core k-mer -- the genome-side index of the off-target search, built once per (PAM, core
    // length) and kept with the image like bowtie's index files (offtarget.cu)
    uint4 *d_site_slots = nullptr;     // per slot {global position, hi plane | strand << 31, lo plane, core key}; pad slots: x = ~0
    uint32_t *d_site_blocks = nullptr; // bit-sliced form of the slots: 40 words per 32 slots (join_blocks.cu), or null
    uint32_t *d_site_off = nullptr;    // n_buckets + 1: first slot of each bucket (multiple of 32) | pad count in the low 5 bits
    uint32_t *d_tickets = nullptr;     // 64 work counters of the block join (one per launch in flight)
    int64_t site_cap = 0;       // slots allocated
    int64_t site_slots = 0;     // slots in use (buckets padded to whole blocks of 32)
    int64_t site_n = -1;        // windows tabled (-1: no table)
    int bj_grid = 0;            // resident-CTA grid of ot_join_blocks_kernel (cached)
    uint8_t site_fw[8] = {}, site_rv[8] = {};  // the PAM sets the table was built for
    int site_pam_len = 0, site_core_len = 0;
    float site_build_ms = 0;
    // Partition of the table by core-key range (multi-GPU off-target search): this handle tables --
    // and therefore searches -- only the windows whose core key has its top log2(n_parts) bits equal
    // to part_ix.  `site_part_*` is what the caller asked for, `site_tab_part_*` what the cached
    // table was built with.
    int site_part_ix = 0, site_n_parts = 1;
    int site_tab_part_ix = 0, site_tab_n_parts = 1;
    uint32_t site_row_off[33] = {};  // first table slot of each 1/32 of the core-key space (+ the slot count)
    // second stream + events: the guide index is built slice by slice while the join already
    // works on the slices that are ready (launch_ot_join)
    cudaStream_t build_stream = nullptr, join_stream = nullptr;
    cudaEvent_t ev_slice[36] = {};  // [0..33]: slices of the overlapped schedule; [34], [35]: the extras of the gathered index
};

int device_init(DeviceState *ds, int device);
void kev_begin(DeviceState *ds, cudaStream_t stream);
void kev_end(DeviceState *ds, cudaStream_t stream);
int ensure_bytes(void **p, int64_t *cap, int64_t need);
int ensure_pinned(DeviceState *ds, int64_t need);

// launch description of one scan over global PAM-start positions [lo, hi)
struct ScanRange {
    uint32_t lo, hi;     // PAM start p must satisfy lo <= p < hi
    uint32_t limit;      // and p + P <= limit
    uint32_t out_base;   // subtracted from reported positions
};

struct PlanesView {
    const ulonglong2 *planes;  // planes[b - blk_lo], valid for blk_lo <= b < blk_lo + blk_n
    int64_t blk_lo;
    int64_t blk_n;
    const uint32_t *run_start, *run_end;
    const uint8_t *run_kind;
    const uint32_t *tile_first_run;
    int n_runs;
};

int launch_pam_scan(DeviceState *ds, const PlanesView &pv, const PamSpec &pam, int mode,
                    const ScanRange &r, uint32_t *d_fw, uint64_t cap_fw, uint32_t *d_rv,
                    uint64_t cap_rv, unsigned long long *d_counts, cudaStream_t stream,
                    int *n_launches);

struct OtParams {
    PamSpec pam;
    uint32_t core_mask;   // compare positions counted against core_max (incl. fixed PAM in raw mode)
    uint32_t cmp_mask;    // compare positions at all
    int32_t core_max, total_max;
    int32_t core_len;
    int32_t raw_mode;     // 0: PAM-exact sites -> hit records; 1: every window -> per-guide flags
};

int launch_ot_scan(DeviceState *ds, const PlanesView &pv, const OtParams &op, const ScanRange &r,
                   const uint2 *d_guides, int64_t n_guides, guido_ot_hit *d_hits, uint64_t cap,
                   unsigned long long *d_count, uint8_t *d_flags, unsigned long long *d_site_count,
                   cudaStream_t stream, int *n_launches);

int exclusive_scan_u32(uint32_t *data, uint32_t n, uint32_t *sums, uint32_t *cursor, cudaStream_t stream);
int sort_hits_device(DeviceState *ds, guido_ot_hit *d_hits, uint64_t n, int64_t n_guides, cudaStream_t stream,
                     int *n_launches);
int encode_guides_device(DeviceState *ds, const char *d_letters, int64_t n, uint2 *d_out, unsigned long long *d_bad,
                         cudaStream_t stream);
int pack_hits_device(DeviceState *ds, const guido_ot_hit *d_hits, uint64_t n, unsigned long long *d_out, cudaStream_t stream);
int concat_hits_device(DeviceState *ds, const guido_ot_hit *d_gathered, int n_parts, uint64_t stride, guido_ot_hit *d_out,
                       uint64_t cap, unsigned long long *d_total, cudaStream_t stream);
int csr_hits_device(DeviceState *ds, const guido_ot_hit *d_hits, uint64_t n, uint32_t n_guides, uint32_t *d_gpos, uint32_t *d_strand,
                    uint32_t *d_offsets, cudaStream_t stream);
int pack_send_device(DeviceState *ds, const guido_ot_hit *d_hits, const unsigned long long *d_count, uint64_t cap,
                     unsigned long long *d_send, cudaStream_t stream);
int concat_packed_device(DeviceState *ds, const unsigned long long *d_gathered, int n_parts, uint64_t stride,
                         unsigned long long *d_out, uint64_t cap, unsigned long long *d_total, cudaStream_t stream);
int mark_hit_guides(DeviceState *ds, const guido_ot_hit *d_hits, uint64_t n, uint8_t *d_flags, cudaStream_t stream);
bool ot_index_applicable(const OtParams &op, int64_t n_guides);

int launch_ot_index(DeviceState *ds, const PlanesView &pv, const OtParams &op, const ScanRange &r,
                    const uint2 *d_guides, int64_t n_guides, guido_ot_hit *d_hits, uint64_t cap,
                    unsigned long long *d_count, unsigned long long *d_site_count,
                    cudaStream_t stream, int *n_launches);

// genome-side PAM-site table (built on first use, cached in DeviceState) + guide-side seed index:
// the search is a bucket-by-bucket join of the two
int ensure_site_table(DeviceState *ds, const PlanesView &pv, const OtParams &op, const ScanRange &r,
                      cudaStream_t stream, bool *built);
void release_site_table(DeviceState *ds);
// Optional: the handle's key range is cut into n equal slices that are indexed and joined one after the
// other; slice s writes its hits to d_hits + s * cap (cap records each), its count to d_count[s], and
// events[s] is recorded on the stream behind its join -- a caller can start sending slice s while the
// later slices are still being searched.
struct OtSlices {
    int n = 1;
    cudaEvent_t *events = nullptr;
    // optional: behind the join of slice s its hits are packed into the 8-byte wire format at
    // send + s * send_stride words (header word = count, then at most send_stride - 1 records), BEFORE
    // events[s] is recorded: what remains for the caller's second stream are copies
    unsigned long long *send = nullptr;
    uint64_t send_stride = 0;
};
int launch_ot_join(DeviceState *ds, const PlanesView &pv, const OtParams &op, const ScanRange &r,
                   const uint2 *d_guides, int64_t n_guides, guido_ot_hit *d_hits, uint64_t cap,
                   unsigned long long *d_count, unsigned long long *d_site_count,
                   cudaStream_t stream, int *n_launches, const OtSlices *slices = nullptr);

}  // namespace guido
