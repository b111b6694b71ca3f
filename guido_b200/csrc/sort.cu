// sort.cu -- canonical order of the off-target hit list: (guide, gpos, strand), on the device.
//
// The search kernels emit hits in completion order (atomic cursor).  The host-buffer C-ABI call
// promises a deterministic order, so 1/2/4/8-GPU runs and repeated runs are byte-identical.
//
// The order has structure a general sort does not use: the major key is the guide (a small dense
// range) and a guide owns ~100 hits.  So: histogram of the guide field, exclusive scan, one scatter
// of the 16-byte records into per-guide segments (a counting sort: every record moves ONCE), then a
// rank sort inside each segment -- one warp per guide, keys (gpos << 1 | strand) broadcast from
// shared memory, rank = number of smaller keys (keys of one guide are distinct), record written to
// its final place.  ~0.4 ms for the 10.8 M hits of the 100 k-guide search where seven radix passes
// over (key, value) pairs took 1.0 ms.  Lists with a guide that owns more than kSegMax hits (highly
// repetitive targets) take the general path: CUB DeviceRadixSort on key = guide << 33 | gpos << 1 |
// strand, value = {hi, lo} (a library call).
#include <cub/device/device_radix_sort.cuh>

#include "device.cuh"

namespace guido {

__global__ void __launch_bounds__(256) split_hits_kernel(const guido_ot_hit *hits, uint64_t n,
                                                          unsigned long long *keys, unsigned long long *vals) {
    for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256ull) {
        const guido_ot_hit h = hits[i];
        keys[i] = ((unsigned long long)h.guide << 33) | ((unsigned long long)h.gpos << 1) | (h.hi >> 31);
        vals[i] = ((unsigned long long)h.hi << 32) | h.lo;
    }
}

__global__ void __launch_bounds__(256) merge_hits_kernel(const unsigned long long *keys, const unsigned long long *vals,
                                                          uint64_t n, guido_ot_hit *hits) {
    for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256ull) {
        const unsigned long long k = keys[i], v = vals[i];
        guido_ot_hit h;
        h.guide = (uint32_t)(k >> 33);
        h.gpos = (uint32_t)((k >> 1) & 0xffffffffull);
        h.hi = (uint32_t)(v >> 32);
        h.lo = (uint32_t)v;
        hits[i] = h;
    }
}

// 16-byte records -> 8-byte wire format: guide << 33 | gpos << 1 | strand (the sort key itself; the
// window planes can be re-derived from the image, guido_hits_expand)
__global__ void __launch_bounds__(256) pack_hits_kernel(const guido_ot_hit *hits, uint64_t n, unsigned long long *out) {
    const uint4 *src = reinterpret_cast<const uint4 *>(hits);
    for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256ull) {
        const uint4 h = src[i];
        out[i] = ((unsigned long long)h.x << 33) | ((unsigned long long)h.y << 1) | (h.z >> 31);
    }
}

int pack_hits_device(DeviceState *ds, const guido_ot_hit *d_hits, uint64_t n, unsigned long long *d_out, cudaStream_t stream) {
    if (n == 0) return GUIDO_OK;
    const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)ds->sm_count * 16);
    pack_hits_kernel<<<grid, 256, 0, stream>>>(d_hits, n, d_out);
    GUIDO_CUDA(cudaGetLastError());
    return GUIDO_OK;
}

// Merge of per-rank hit lists after ONE padded all-gather: part r sits at gathered[r * stride ..],
// its record 0 is a header whose first 8 bytes hold the part's hit count, the hits follow.  The parts
// are copied back to back into `out`; *total receives the sum.  No host round trip between the
// search, the gather and this kernel.
__global__ void __launch_bounds__(256) concat_hits_kernel(const uint4 *gathered, int n_parts, unsigned long long stride,
                                                           uint4 *out, unsigned long long cap, unsigned long long *total) {
    __shared__ unsigned long long s_off[65];
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int r = 0; r < n_parts; r++) {
            const uint4 h = gathered[(size_t)r * stride];
            unsigned long long c = ((unsigned long long)h.y << 32) | h.x;
            if (c > stride - 1) c = stride - 1;  // a part that overflowed its slot: the caller re-runs with a larger stride
            s_off[r] = run;
            run += c;
        }
        s_off[n_parts] = run;
        if (blockIdx.x == 0) *total = run;
    }
    __syncthreads();
    for (int r = 0; r < n_parts; r++) {
        const unsigned long long n = s_off[r + 1] - s_off[r];
        const uint4 *src = gathered + (size_t)r * stride + 1;
        for (unsigned long long i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * 256ull)
            if (s_off[r] + i < cap) out[s_off[r] + i] = src[i];
    }
}

int concat_hits_device(DeviceState *ds, const guido_ot_hit *d_gathered, int n_parts, uint64_t stride, guido_ot_hit *d_out,
                       uint64_t cap, unsigned long long *d_total, cudaStream_t stream) {
    if (n_parts < 1 || n_parts > 64 || stride < 1) { set_error("bad part count or stride"); return GUIDO_ERR_ARG; }
    concat_hits_kernel<<<ds->sm_count * 4, 256, 0, stream>>>(reinterpret_cast<const uint4 *>(d_gathered), n_parts, stride,
                                                            reinterpret_cast<uint4 *>(d_out), cap, d_total);
    GUIDO_CUDA(cudaGetLastError());
    return GUIDO_OK;
}

// Sorted 16-byte records -> CSR by guide: positions (4 bytes per hit), strand bits (1 bit per hit) and
// the first hit of every guide.  4.1 bytes per hit on the wire instead of 8 or 16; nothing is lost
// (the guide of hit i is the segment i falls in, the window planes come from the image).
__global__ void __launch_bounds__(256) csr_hits_kernel(const guido_ot_hit *hits, uint64_t n, uint32_t n_guides, uint32_t *gpos,
                                                        uint32_t *strand_bits, uint32_t *offsets) {
    const uint4 *src = reinterpret_cast<const uint4 *>(hits);
    const uint64_t n_round = (n + 31) & ~31ull;  // whole warps: the strand word of the last 32 hits is a ballot too
    for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n_round; i += (uint64_t)gridDim.x * 256ull) {
        const bool in = i < n;
        const uint4 h = in ? src[i] : make_uint4(0, 0, 0, 0);
        const uint32_t bits = __ballot_sync(0xffffffffu, in && (h.z >> 31));
        if ((threadIdx.x & 31) == 0) strand_bits[i >> 5] = bits;
        if (in) {
            gpos[i] = h.y;
            // guides (previous hit's guide, this guide] start here
            const uint32_t prev = i ? src[i - 1].x + 1u : 0u;
            for (uint32_t g = prev; g <= h.x; g++) offsets[g] = (uint32_t)i;
            if (i == n - 1) for (uint32_t g = h.x + 1u; g <= n_guides; g++) offsets[g] = (uint32_t)n;
        }
    }
    if (n == 0) for (uint32_t g = blockIdx.x * 256u + threadIdx.x; g <= n_guides; g += gridDim.x * 256u) offsets[g] = 0u;
}

int csr_hits_device(DeviceState *ds, const guido_ot_hit *d_hits, uint64_t n, uint32_t n_guides, uint32_t *d_gpos, uint32_t *d_strand,
                    uint32_t *d_offsets, cudaStream_t stream) {
    const int grid = (int)std::min<uint64_t>(std::max<uint64_t>((n + 255) / 256, 1), (uint64_t)ds->sm_count * 16);
    csr_hits_kernel<<<grid, 256, 0, stream>>>(d_hits, n, n_guides, d_gpos, d_strand, d_offsets);
    GUIDO_CUDA(cudaGetLastError());
    return GUIDO_OK;
}

// The same merge on the 8-byte wire format (half the NVLink bytes): a rank packs its hits behind a
// header word that holds its count -- the count is read on the device, so the step needs no host
// round trip -- and the gathered parts are copied back to back.
__global__ void __launch_bounds__(256) pack_send_kernel(const guido_ot_hit *hits, const unsigned long long *count, uint64_t cap,
                                                         unsigned long long *send) {
    const unsigned long long n = min(*count, (unsigned long long)cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) send[0] = *count;
    const uint4 *src = reinterpret_cast<const uint4 *>(hits);
    for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256ull) {
        const uint4 h = src[i];
        send[1 + i] = ((unsigned long long)h.x << 33) | ((unsigned long long)h.y << 1) | (h.z >> 31);
    }
}

__global__ void __launch_bounds__(256) concat_packed_kernel(const unsigned long long *gathered, int n_parts, unsigned long long stride,
                                                             unsigned long long *out, unsigned long long cap, unsigned long long *total) {
    __shared__ unsigned long long s_off[65];
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int r = 0; r < n_parts; r++) {
            unsigned long long c = gathered[(size_t)r * stride];
            if (c > stride - 1) c = stride - 1;  // a part that overflowed its slot: the caller re-runs with a larger stride
            s_off[r] = run;
            run += c;
        }
        s_off[n_parts] = run;
        if (blockIdx.x == 0) *total = run;
    }
    __syncthreads();
    for (int r = 0; r < n_parts; r++) {
        const unsigned long long n = s_off[r + 1] - s_off[r];
        const unsigned long long *src = gathered + (size_t)r * stride + 1;
        for (unsigned long long i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * 256ull)
            if (s_off[r] + i < cap) out[s_off[r] + i] = src[i];
    }
}

int pack_send_device(DeviceState *ds, const guido_ot_hit *d_hits, const unsigned long long *d_count, uint64_t cap,
                     unsigned long long *d_send, cudaStream_t stream) {
    pack_send_kernel<<<ds->sm_count * 8, 256, 0, stream>>>(d_hits, d_count, cap, d_send);
    GUIDO_CUDA(cudaGetLastError());
    return GUIDO_OK;
}

int concat_packed_device(DeviceState *ds, const unsigned long long *d_gathered, int n_parts, uint64_t stride,
                         unsigned long long *d_out, uint64_t cap, unsigned long long *d_total, cudaStream_t stream) {
    if (n_parts < 1 || n_parts > 64 || stride < 1) { set_error("bad part count or stride"); return GUIDO_ERR_ARG; }
    concat_packed_kernel<<<ds->sm_count * 4, 256, 0, stream>>>(d_gathered, n_parts, stride, d_out, cap, d_total);
    GUIDO_CUDA(cudaGetLastError());
    return GUIDO_OK;
}

__global__ void __launch_bounds__(256) mark_guides_kernel(const guido_ot_hit *hits, uint64_t n, uint8_t *flags) {
    for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256ull)
        flags[hits[i].guide] = 1;
}

// flags[g] = 1 for every guide that owns at least one hit (flags must be zeroed by the caller)
int mark_hit_guides(DeviceState *ds, const guido_ot_hit *d_hits, uint64_t n, uint8_t *d_flags, cudaStream_t stream) {
    if (n == 0) return GUIDO_OK;
    const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)ds->sm_count * 16);
    mark_guides_kernel<<<grid, 256, 0, stream>>>(d_hits, n, d_flags);
    GUIDO_CUDA(cudaGetLastError());
    return GUIDO_OK;
}

// n x 20 letters -> n x {hi, lo} plane words; *bad = smallest (guide * 32 + position) holding a
// letter outside A/C/G/T (left untouched when every letter is fine)
__global__ void __launch_bounds__(256) encode_guides_kernel(const char *letters, uint32_t n, uint2 *out,
                                                             unsigned long long *bad) {
    const uint32_t g = blockIdx.x * 256u + threadIdx.x;
    if (g >= n) return;
    const char *p = letters + (size_t)g * kProto;
    uint32_t hi = 0, lo = 0;
#pragma unroll
    for (int i = 0; i < kProto; i++) {
        const uint32_t c = (uint32_t)(unsigned char)p[i] & 0xDFu;  // upper case
        // A 0x41, C 0x43, G 0x47, T 0x54: hi plane = bit 2, lo plane = bit 1 ^ bit 2
        if (c != 0x41 && c != 0x43 && c != 0x47 && c != 0x54) atomicMin(bad, (unsigned long long)g * 32ull + i);
        hi |= ((c >> 2) & 1u) << i;
        lo |= (((c >> 1) ^ (c >> 2)) & 1u) << i;
    }
    out[g] = make_uint2(hi, lo);
}

int encode_guides_device(DeviceState *ds, const char *d_letters, int64_t n, uint2 *d_out, unsigned long long *d_bad,
                         cudaStream_t stream) {
    if (n == 0) return GUIDO_OK;
    encode_guides_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_letters, (uint32_t)n, d_out, d_bad);
    GUIDO_CUDA(cudaGetLastError());
    return GUIDO_OK;
}

// ---- counting sort by guide + rank sort per guide ------------------------------------------------

constexpr uint32_t kSegMax = 4096;   // largest per-guide segment the rank sort takes
constexpr uint32_t kSegSmem = 512;   // ... with its keys in shared memory (larger ones read them from L1/L2)
constexpr int kSegWarps = 8;

__global__ void __launch_bounds__(256) hit_hist_kernel(const guido_ot_hit *hits, uint64_t n, uint32_t n_guides,
                                                        uint32_t *counts) {
    for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256ull)
        atomicAdd(&counts[min(hits[i].guide, n_guides - 1u)], 1u);  // (a guide field out of range stays memory safe)
}

__global__ void __launch_bounds__(256) seg_max_kernel(const uint32_t *counts, uint32_t n_guides, uint32_t *out_max) {
    uint32_t m = 0;
    for (uint32_t g = blockIdx.x * 256u + threadIdx.x; g < n_guides; g += gridDim.x * 256u) m = max(m, counts[g]);
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out_max, m);
}

__global__ void __launch_bounds__(256) hit_scatter_kernel(const guido_ot_hit *hits, uint64_t n, uint32_t n_guides,
                                                           uint32_t *cursor, guido_ot_hit *out) {
    const uint4 *src = reinterpret_cast<const uint4 *>(hits);
    uint4 *dst = reinterpret_cast<uint4 *>(out);
    for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256ull) {
        const uint4 h = src[i];  // {guide, gpos, hi, lo}
        dst[atomicAdd(&cursor[min(h.x, n_guides - 1u)], 1u)] = h;
    }
}

// One warp per guide.  seg = the guide's records in arrival order, out = their final places.
__global__ void __launch_bounds__(kSegWarps * 32) seg_rank_sort_kernel(const guido_ot_hit *seg, const uint32_t *offsets,
                                                                       uint32_t n_guides, guido_ot_hit *out) {
    __shared__ unsigned long long s_keys[kSegWarps][kSegSmem];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint4 *src = reinterpret_cast<const uint4 *>(seg);
    uint4 *dst = reinterpret_cast<uint4 *>(out);
    unsigned long long *keys = s_keys[warp];
    for (uint32_t g = blockIdx.x * kSegWarps + warp; g < n_guides; g += gridDim.x * kSegWarps) {
        const uint32_t lo = __ldg(&offsets[g]), s = __ldg(&offsets[g + 1]) - lo;
        if (s == 0) continue;
        const bool in_smem = s <= kSegSmem;
        if (in_smem) {
            __syncwarp();  // the previous segment's readers are done
            for (uint32_t j = lane; j < s; j += 32) {
                const uint4 h = src[lo + j];
                keys[j] = ((unsigned long long)h.y << 1) | (h.z >> 31);
            }
            __syncwarp();
        }
        // four records per lane and sweep: every key that is read is compared four times
        for (uint32_t i0 = 0; i0 < s; i0 += 128) {
            uint4 rec[4];
            unsigned long long k[4];
            uint32_t rank[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = i0 + u * 32 + lane;
                rec[u] = i < s ? src[lo + i] : make_uint4(0, 0, 0, 0);
                k[u] = ((unsigned long long)rec[u].y << 1) | (rec[u].z >> 31);
                rank[u] = 0;
            }
            if (in_smem) {
                for (uint32_t j = 0; j < s; j++) {
                    const unsigned long long kj = keys[j];
#pragma unroll
                    for (int u = 0; u < 4; u++) rank[u] += kj < k[u];
                }
            } else {
                for (uint32_t j = 0; j < s; j++) {
                    const uint4 h = __ldg(&src[lo + j]);
                    const unsigned long long kj = ((unsigned long long)h.y << 1) | (h.z >> 31);
#pragma unroll
                    for (int u = 0; u < 4; u++) rank[u] += kj < k[u];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (i0 + u * 32 + lane < s) dst[lo + rank[u]] = rec[u];
        }
    }
}

static int sort_hits_segmented(DeviceState *ds, guido_ot_hit *d_hits, uint64_t n, int64_t n_guides, cudaStream_t stream,
                               int *n_launches, bool *done) {
    *done = false;
    if (n_guides < 1 || n_guides > (1ll << 27) || n > 0xffffffffull) return GUIDO_OK;  // general path
    const uint32_t ng = (uint32_t)n_guides, n_scan = ng + 1, n_scan_blocks = (n_scan + 4095) / 4096;
    const size_t rec_bytes = (n * sizeof(guido_ot_hit) + 255) & ~(size_t)255;
    const size_t cnt_bytes = ((size_t)n_scan * 4 + 255) & ~(size_t)255;
    const size_t sum_bytes = ((size_t)n_scan_blocks * 4 + 255) & ~(size_t)255;
    int rc = ensure_bytes(&ds->d_out[1], &ds->out_cap[1], (int64_t)(rec_bytes + 2 * cnt_bytes + sum_bytes + 256));
    if (rc) return rc;
    char *base = (char *)ds->d_out[1];
    guido_ot_hit *tmp = (guido_ot_hit *)base;
    uint32_t *counts = (uint32_t *)(base + rec_bytes), *cursor = (uint32_t *)(base + rec_bytes + cnt_bytes);
    uint32_t *sums = (uint32_t *)(base + rec_bytes + 2 * cnt_bytes), *d_max = (uint32_t *)(base + rec_bytes + 2 * cnt_bytes + sum_bytes);
    const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)ds->sm_count * 16);
    GUIDO_CUDA(cudaMemsetAsync(counts, 0, cnt_bytes, stream));
    GUIDO_CUDA(cudaMemsetAsync(d_max, 0, 4, stream));
    hit_hist_kernel<<<grid, 256, 0, stream>>>(d_hits, n, ng, counts);
    seg_max_kernel<<<(int)std::min<uint32_t>((ng + 255) / 256, (uint32_t)ds->sm_count * 4), 256, 0, stream>>>(counts, ng, d_max);
    uint32_t seg_max = 0;
    GUIDO_CUDA(cudaMemcpyAsync(&seg_max, d_max, 4, cudaMemcpyDeviceToHost, stream));
    GUIDO_CUDA(cudaStreamSynchronize(stream));  // 4 bytes: which path
    if (n_launches) *n_launches += 2;
    if (seg_max > kSegMax) return GUIDO_OK;  // a guide with a very long list: general path
    if ((rc = exclusive_scan_u32(counts, n_scan, sums, cursor, stream))) return rc;
    hit_scatter_kernel<<<grid, 256, 0, stream>>>(d_hits, n, ng, cursor, tmp);
    const int sgrid = (int)std::min<uint32_t>((ng + kSegWarps - 1) / kSegWarps, (uint32_t)ds->sm_count * 32);
    seg_rank_sort_kernel<<<sgrid, kSegWarps * 32, 0, stream>>>(tmp, counts, ng, d_hits);
    GUIDO_CUDA(cudaGetLastError());
    if (n_launches) *n_launches += 5;
    *done = true;
    return GUIDO_OK;
}

// sort d_hits[0..n) in place; scratch comes from ds->d_out[1]
int sort_hits_device(DeviceState *ds, guido_ot_hit *d_hits, uint64_t n, int64_t n_guides, cudaStream_t stream,
                     int *n_launches) {
    if (n < 2) return GUIDO_OK;
    const char *force = getenv("GUIDO_SORT");  // "radix": always the general path (tests, comparisons)
    if (!(force && force[0] == 'r')) {
        bool done = false;
        int rc = sort_hits_segmented(ds, d_hits, n, n_guides, stream, n_launches, &done);
        if (rc || done) return rc;
    }
    int guide_bits = 1;
    while ((1ll << guide_bits) < n_guides && guide_bits < 31) guide_bits++;
    const int end_bit = 33 + guide_bits;
    size_t temp_bytes = 0;
    cub::DoubleBuffer<unsigned long long> kb(nullptr, nullptr), vb(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, kb, vb, (int64_t)n, 0, end_bit, stream);
    const size_t arr = (n * 8 + 255) & ~(size_t)255;
    int rc = ensure_bytes(&ds->d_out[1], &ds->out_cap[1], (int64_t)(4 * arr + temp_bytes + 256));
    if (rc) return rc;
    char *base = (char *)ds->d_out[1];
    unsigned long long *k0 = (unsigned long long *)base, *k1 = (unsigned long long *)(base + arr),
                       *v0 = (unsigned long long *)(base + 2 * arr), *v1 = (unsigned long long *)(base + 3 * arr);
    void *temp = base + 4 * arr;
    const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)ds->sm_count * 16);
    split_hits_kernel<<<grid, 256, 0, stream>>>(d_hits, n, k0, v0);
    kb = cub::DoubleBuffer<unsigned long long>(k0, k1);
    vb = cub::DoubleBuffer<unsigned long long>(v0, v1);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, kb, vb, (int64_t)n, 0, end_bit, stream);
    if (e != cudaSuccess) { set_error("radix sort failed: %s", cudaGetErrorString(e)); return GUIDO_ERR_CUDA; }
    merge_hits_kernel<<<grid, 256, 0, stream>>>(kb.Current(), vb.Current(), n, d_hits);
    GUIDO_CUDA(cudaGetLastError());
    if (n_launches) *n_launches += 2 + (end_bit + 7) / 8;
    return GUIDO_OK;
}

}  // namespace guido
