// scan_common.cuh -- device helpers shared by the PAM-scan and off-target kernels.
//
// Genome layout in HBM: one ulonglong2 {x = lo plane, y = hi plane} per 64 bases, bit i of a word
// <-> base 64*b + i.  One thread owns one 64-base block; the halo (previous / next block) comes
// from the neighbouring lane by warp shuffle, the warp-edge lanes load it themselves.
#pragma once
#include "device.cuh"

namespace guido {

constexpr int kThreads = 256;               // threads per CTA in the scan kernels
constexpr int kTileBases = kThreads * 64;   // 16384 bases per tile
constexpr unsigned kFull = 0xffffffffu;
constexpr int kTileRunShift = 16;          // tile_first_run has one entry per 65536 bases
constexpr int kScanRawWindow = 3;         // internal: both orientations own window [p, p+P+20)

// positions whose base belongs to set s (bit0 A, bit1 C, bit2 G, bit3 T); s is warp-uniform
__device__ __forceinline__ uint64_t member_mask(uint64_t lo, uint64_t hi, uint32_t s) {
    uint64_t r = 0;
    if (s & 1) r |= ~hi & ~lo;
    if (s & 2) r |= ~hi & lo;
    if (s & 4) r |= hi & ~lo;
    if (s & 8) r |= hi & lo;
    return r;
}

// bit i set <=> the PAM (sets[0..P-1]) matches at base 64*b + i; needs the next block for the tail
__device__ __forceinline__ uint64_t pam_hit_mask(ulonglong2 cur, ulonglong2 nxt, const uint8_t *sets,
                                                 int P) {
    uint64_t h = ~0ull;
    for (int i = 0; i < P; i++) {
        uint32_t s = sets[i];
        if (s == 15) continue;  // N: every base; non-ACGT bases are removed by the run veto
        uint64_t mc = member_mask(cur.x, cur.y, s);
        if (i) {
            uint64_t mn = member_mask(nxt.x, nxt.y, s);
            mc = (mc >> i) | (mn << (64 - i));
        }
        h &= mc;
    }
    return h;
}

// bits [lo_bit, hi_bit) of a 64-bit mask
__device__ __forceinline__ uint64_t bit_range(int lo_bit, int hi_bit) {
    if (hi_bit <= lo_bit) return 0;
    uint64_t a = (hi_bit >= 64) ? ~0ull : ((1ull << hi_bit) - 1);
    uint64_t b = (lo_bit <= 0) ? 0 : ((1ull << lo_bit) - 1);
    return a & ~b;
}

// restrict a block's hit mask to PAM starts p with r.lo <= p < r.hi and p + P <= r.limit
__device__ __forceinline__ uint64_t range_mask(int64_t block, const ScanRange &r, int P) {
    int64_t base = block * 64;
    int64_t hi = (int64_t)r.hi;
    int64_t lim = (int64_t)r.limit - P + 1;
    if (lim < hi) hi = lim;
    int64_t a = (int64_t)r.lo - base, b = hi - base;
    if (a < 0) a = 0;
    if (b > 64) b = 64;
    return bit_range((int)a, (int)b);
}

// first run whose end is > pos (runs are sorted and disjoint, so ends are sorted too)
__device__ __forceinline__ int first_run_ending_after(const PlanesView &pv, uint32_t pos) {
    int lo = 0, hi = pv.n_runs;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (pv.run_end[mid] > pos) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// Run range [r_lo, r_hi) touching positions [t0 - 32, t1 + 32), and whether one run covers the
// whole [t0, t1).  pv.tile_first_run[t0 >> 16] (built at upload) gives the starting point, so the
// hot path does one table load instead of a binary search; every thread computes it (uniform,
// cached loads) and no shared-memory broadcast / barrier is needed.
__device__ __forceinline__ void tile_runs(const PlanesView &pv, uint32_t t0, uint32_t t1, int &r_lo, int &r_hi,
                                          bool &skip) {
    const uint32_t entry = __ldg(&pv.tile_first_run[t0 >> kTileRunShift]);
    if (!(entry >> 31)) {  // no run anywhere near this 65536-base tile: the common case
        r_lo = r_hi = 0;
        skip = false;
        return;
    }
    int r = (int)(entry & 0x7fffffffu);
    const uint32_t from = t0 >= 32 ? t0 - 32 : 0;
    while (r < pv.n_runs && __ldg(&pv.run_end[r]) <= from) r++;
    r_lo = r;
    skip = false;
    while (r < pv.n_runs && __ldg(&pv.run_start[r]) < t1 + 32) {
        if (__ldg(&pv.run_start[r]) <= t0 && __ldg(&pv.run_end[r]) >= t1) skip = true;
        r++;
    }
    r_hi = r;
}

__device__ __forceinline__ bool span_touches_run(const PlanesView &pv, int r_lo, int r_hi, uint32_t a,
                                                 uint32_t b, bool kind0_only) {
    for (int r = r_lo; r < r_hi; r++) {
        if (pv.run_start[r] < b && pv.run_end[r] > a && (!kind0_only || pv.run_kind[r] == 0))
            return true;
    }
    return false;
}

// Remove from (hf, hr) the PAM starts that the scan mode rejects.  Only called in tiles that
// overlap at least one run.  fw window = [p-20, p+P), rv window = [p, p+P+20).
__device__ __forceinline__ void veto_hits(const PlanesView &pv, int r_lo, int r_hi, int mode, int P,
                                          uint32_t base, uint64_t &hf, uint64_t &hr) {
    uint64_t m = hf;
    while (m) {
        int i = __ffsll((long long)m) - 1;
        m &= m - 1;
        uint32_t p = base + i;
        bool bad = span_touches_run(pv, r_lo, r_hi, p, p + P, false);
        if (!bad && mode == GUIDO_SCAN_GUIDE) bad = span_touches_run(pv, r_lo, r_hi, p - kProto, p + P, true);
        if (!bad && mode == GUIDO_SCAN_STRICT) bad = span_touches_run(pv, r_lo, r_hi, p - kProto, p + P, false);
        if (!bad && mode == kScanRawWindow) bad = span_touches_run(pv, r_lo, r_hi, p, p + P + kProto, false);
        if (bad) hf &= ~(1ull << i);
    }
    m = hr;
    while (m) {
        int i = __ffsll((long long)m) - 1;
        m &= m - 1;
        uint32_t p = base + i;
        uint32_t e = (mode == GUIDO_SCAN_REGEX) ? p + P : p + P + kProto;
        if (span_touches_run(pv, r_lo, r_hi, p, e, false)) hr &= ~(1ull << i);
    }
}

// bounds-checked block load: blocks outside the resident range read as zeros (their positions
// are always masked out by the scan range, so the value never matters)
__device__ __forceinline__ ulonglong2 load_block(const PlanesView &pv, int64_t block) {
    int64_t i = block - pv.blk_lo;
    if (i < 0 || i >= pv.blk_n) return make_ulonglong2(0ull, 0ull);
    return pv.planes[i];
}

__device__ __forceinline__ ulonglong2 shfl_down_u2(ulonglong2 v, int d) {
    ulonglong2 r;
    r.x = __shfl_down_sync(kFull, v.x, d);
    r.y = __shfl_down_sync(kFull, v.y, d);
    return r;
}
__device__ __forceinline__ ulonglong2 shfl_up_u2(ulonglong2 v, int d) {
    ulonglong2 r;
    r.x = __shfl_up_sync(kFull, v.x, d);
    r.y = __shfl_up_sync(kFull, v.y, d);
    return r;
}

// exclusive scan of a packed pair of 16-bit counters over the CTA; returns the exclusive prefix,
// *total receives the CTA aggregate.  s_warp: kThreads/32 words of shared memory.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *s_warp, uint32_t *total) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(kFull, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t warp_prefix = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) {
        uint32_t x = s_warp[w];
        if (w < warp) warp_prefix += x;
        tot += x;
    }
    *total = tot;
    __syncthreads();  // s_warp may be reused by the caller
    return warp_prefix + inc - v;
}

}  // namespace guido
