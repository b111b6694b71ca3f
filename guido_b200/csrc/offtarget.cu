// offtarget.cu -- batched <=k-mismatch off-target search over the bit-plane genome image.
//
// Sits under run_bowtie (reference guido/off_targets.py:89-227): replaces the bowtie 1
// `-n/-l/-e -a -y` subprocess.  Two kernels:
//
//   ot_scan_kernel   brute force: every PAM-valid window (both strands) x every guide, by plane
//                    XOR + OR + popc.  Integer-issue bound.  Also the raw-alignment flag pass
//                    (all windows, PAM letters inside the compare) that decides key presence.
//   ot_index_kernel  (offtarget_index.cu) guide-side seed index; windows look up their core.
//
// Plane trick: with guides and windows both stored as two 1-bit planes (hi, lo) in GUIDE
// orientation, the per-position mismatch mask of all 20 positions is two LOP3s:
//     m = (A_w ^ A_g) | (B_w ^ B_g)
// and the mismatch count is one POPC -- no 2-bit fold, no shifts in the inner loop.
#include <algorithm>

#include "scan_common.cuh"

namespace guido {

constexpr int kRing = 4096;        // site ring buffer entries per CTA (3 x u32 each = 48 KB)
constexpr int kSitesPerThread = 8;
constexpr int kRound = kThreads * kSitesPerThread;  // 2048 sites compared per round

struct OtScanArgs {
    PlanesView pv;
    OtParams op;
    ScanRange r;
    int64_t first_block;
    uint32_t n_tiles;
    const uint2 *guides;  // x = hi plane, y = lo plane, pre-masked with cmp_mask
    uint32_t n_guides;
    guido_ot_hit *hits;
    unsigned long long cap;
    unsigned long long *count;
    uint8_t *flags;
    unsigned long long *site_count;
};

// W bits starting at bit s of the 192-bit string w0:w1:w2 (s < 128)
__device__ __forceinline__ uint32_t extract_bits(uint64_t w0, uint64_t w1, uint64_t w2, int s, uint32_t wmask) {
    uint64_t a = (s < 64) ? w0 : w1, b = (s < 64) ? w1 : w2;
    int sh = s & 63;
    uint64_t v = a >> sh;
    if (sh) v |= b << (64 - sh);
    return (uint32_t)v & wmask;
}

// compare one round of sites held in the ring [head, head+n) against every guide
__device__ __forceinline__ void compare_round(const OtScanArgs &a, const uint32_t *sA, const uint32_t *sB,
                                              const uint32_t *sQ, uint32_t head, uint32_t n) {
    uint32_t A[kSitesPerThread], B[kSitesPerThread];
    uint32_t valid = 0;
    const uint32_t cmp = a.op.cmp_mask;
#pragma unroll
    for (int k = 0; k < kSitesPerThread; k++) {
        uint32_t j = k * kThreads + threadIdx.x;
        uint32_t idx = (head + j) & (kRing - 1);
        if (j < n) { A[k] = sA[idx] & cmp; B[k] = sB[idx] & cmp; valid |= 1u << k; }
        else { A[k] = 0xffffffffu; B[k] = 0xffffffffu; }
    }
    const int tot_max = a.op.total_max, core_max = a.op.core_max;
    const uint32_t core_mask = a.op.core_mask;
    const uint2 *__restrict__ guides = a.guides;
#pragma unroll 2
    for (uint32_t g = 0; g < a.n_guides; g++) {
        const uint2 G = __ldg(&guides[g]);
        bool any = false;
#pragma unroll
        for (int k = 0; k < kSitesPerThread; k++) {
            uint32_t m = (A[k] ^ G.x) | (B[k] ^ G.y);
            any |= (__popc(m) <= tot_max);
        }
        if (any) {  // rare: ~4e-7 per pair for <=4 of 20
#pragma unroll
            for (int k = 0; k < kSitesPerThread; k++) {
                uint32_t m = (A[k] ^ G.x) | (B[k] ^ G.y);
                if (!((valid >> k) & 1) || __popc(m) > tot_max || __popc(m & core_mask) > core_max) continue;
                if (a.op.raw_mode) {
                    a.flags[g] = 1;
                } else {
                    unsigned long long slot = atomicAdd(a.count, 1ull);
                    if (slot < a.cap) {
                        uint32_t idx = (head + k * kThreads + threadIdx.x) & (kRing - 1);
                        guido_ot_hit h;
                        h.guide = g; h.gpos = sQ[idx]; h.hi = sA[idx]; h.lo = sB[idx];
                        a.hits[slot] = h;
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads) ot_scan_kernel(const OtScanArgs a) {
    extern __shared__ uint32_t s_dyn[];
    uint32_t *sA = s_dyn, *sB = s_dyn + kRing, *sQ = s_dyn + 2 * kRing;
    __shared__ uint32_t s_warp[kThreads / 32];
    __shared__ int s_run[3];
    const int lane = threadIdx.x & 31;
    const int P = a.op.pam.len;
    const int W = kProto + P;
    const uint32_t wmask = (1u << W) - 1;
    const int raw = a.op.raw_mode;
    uint32_t head = 0, qlen = 0;     // ring state (uniform across the CTA)
    unsigned long long n_sites = 0;

    for (uint32_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int64_t tile_block = a.first_block + (int64_t)tile * kThreads;
        const uint32_t t0 = (uint32_t)(tile_block * 64), t1 = t0 + kTileBases;
        __syncthreads();
        if (threadIdx.x == 0) {
            int r_lo = first_run_ending_after(a.pv, t0 >= 32 ? t0 - 32 : 0);
            int r_hi = r_lo, skip = 0;
            while (r_hi < a.pv.n_runs && a.pv.run_start[r_hi] < t1 + 32) {
                if (a.pv.run_start[r_hi] <= t0 && a.pv.run_end[r_hi] >= t1) skip = 1;
                r_hi++;
            }
            s_run[0] = r_lo; s_run[1] = r_hi; s_run[2] = skip;
        }
        __syncthreads();
        if (s_run[2]) continue;
        const int r_lo = s_run[0], r_hi = s_run[1];
        const int64_t block = tile_block + threadIdx.x;
        const ulonglong2 cur = load_block(a.pv, block);
        ulonglong2 nxt = shfl_down_u2(cur, 1), prv = shfl_up_u2(cur, 1);
        if (lane == 31) nxt = load_block(a.pv, block + 1);
        if (lane == 0) prv = load_block(a.pv, block - 1);
        const uint64_t rm = range_mask(block, a.r, raw ? W : P);
        uint64_t hf, hr;
        if (raw) {
            hf = hr = rm;
        } else {
            hf = pam_hit_mask(cur, nxt, a.op.pam.fw, P) & rm;
            hr = pam_hit_mask(cur, nxt, a.op.pam.rv, P) & rm;
        }
        if (r_hi > r_lo && (hf | hr))
            veto_hits(a.pv, r_lo, r_hi, raw ? kScanRawWindow : GUIDO_SCAN_STRICT, P, (uint32_t)(block * 64), hf, hr);

        // append this tile's sites to the ring; run full compare rounds whenever 2048 are queued
        while (true) {
            uint32_t total;
            const uint32_t mine = __popcll(hf) + __popcll(hr);
            uint32_t off = block_exclusive_scan(mine, s_warp, &total);
            if (total == 0) break;
            const uint32_t space = kRing - qlen;
            const uint32_t base = (uint32_t)(block * 64);
            uint32_t appended = total < space ? total : space;
            // forward-strand sites first, then reverse, lowest bit first; stop at `space`
            for (int strand = 0; strand < 2; strand++) {
                uint64_t m = strand ? hr : hf;
                while (m && off < space) {
                    int i = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    uint32_t hi_bits, lo_bits, q;
                    if (strand == 0) {
                        int s = raw ? (64 + i) : (64 + i - kProto);
                        hi_bits = extract_bits(prv.y, cur.y, nxt.y, s, wmask);
                        lo_bits = extract_bits(prv.x, cur.x, nxt.x, s, wmask);
                        q = raw ? base + i : base + i - kProto;
                    } else {  // reverse strand: guide orientation = reverse complement of the window
                        int s = 64 + i;
                        hi_bits = (__brev(extract_bits(prv.y, cur.y, nxt.y, s, wmask)) >> (32 - W)) ^ wmask;
                        lo_bits = (__brev(extract_bits(prv.x, cur.x, nxt.x, s, wmask)) >> (32 - W)) ^ wmask;
                        hi_bits |= 0x80000000u;
                        q = base + i;
                    }
                    uint32_t idx = (head + qlen + off) & (kRing - 1);
                    sA[idx] = hi_bits; sB[idx] = lo_bits; sQ[idx] = q;
                    off++;
                }
                if (strand == 0) hf = m; else hr = m;
            }
            qlen += appended;
            n_sites += (threadIdx.x == 0) ? appended : 0;
            __syncthreads();
            while (qlen >= kRound) {
                compare_round(a, sA, sB, sQ, head, kRound);
                head = (head + kRound) & (kRing - 1);
                qlen -= kRound;
            }
            __syncthreads();
            if (appended == total) break;
        }
    }
    __syncthreads();
    if (qlen) compare_round(a, sA, sB, sQ, head, qlen);
    if (threadIdx.x == 0 && a.site_count) atomicAdd(a.site_count, n_sites);
}

int launch_ot_scan(DeviceState *ds, const PlanesView &pv, const OtParams &op, const ScanRange &r,
                   const uint2 *d_guides, int64_t n_guides, guido_ot_hit *d_hits, uint64_t cap,
                   unsigned long long *d_count, uint8_t *d_flags, unsigned long long *d_site_count,
                   cudaStream_t stream, int *n_launches) {
    if (r.hi <= r.lo || n_guides == 0) return GUIDO_OK;
    OtScanArgs a;
    a.pv = pv; a.op = op; a.r = r;
    a.first_block = r.lo / 64;
    int64_t last_block = ((int64_t)r.hi - 1) / 64;
    a.n_tiles = (uint32_t)((last_block - a.first_block + 1 + kThreads - 1) / kThreads);
    a.guides = d_guides; a.n_guides = (uint32_t)n_guides;
    a.hits = d_hits; a.cap = cap; a.count = d_count; a.flags = d_flags; a.site_count = d_site_count;
    const int smem = 3 * kRing * (int)sizeof(uint32_t);
    GUIDO_CUDA(cudaFuncSetAttribute(ot_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = 0;
    GUIDO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ot_scan_kernel, kThreads, smem));
    int grid = ds->sm_count * std::max(per_sm, 1);
    if ((uint32_t)grid > a.n_tiles) grid = (int)a.n_tiles;
    ot_scan_kernel<<<grid, kThreads, smem, stream>>>(a);
    GUIDO_CUDA(cudaGetLastError());
    if (n_launches) *n_launches += 1;
    return GUIDO_OK;
}

}  // namespace guido
