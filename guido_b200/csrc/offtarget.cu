// offtarget.cu -- batched <=k-mismatch off-target search over the bit-plane genome image.
//
// Sits under run_bowtie (reference guido/off_targets.py:89-227): replaces the bowtie 1
// `-n/-l/-e -a -y` subprocess.  One kernel template, two ways to consume the PAM-valid windows:
//
//   ot_kernel<false>  brute force: every PAM-valid window (both strands) x every guide, by plane
//                     XOR + OR + popc.  Integer-issue bound.  Also the raw-alignment flag pass
//                     (all windows, PAM letters inside the compare) that decides key presence.
//   ot_kernel<true>   guide-side seed index: a window looks up the bucket of its core k-mer and
//                     compares only the distal bases of the guides filed there.  Memory bound.
//
// Both share the front end: a CTA streams 16384-base tiles, extracts the PAM-valid windows of both
// strands into a shared-memory ring (guide orientation, two bit planes per window) and hands full
// rounds of windows to the consumer, so the consumer always works on dense batches.
//
// Plane trick: with guides and windows both stored as two 1-bit planes (hi, lo) in GUIDE
// orientation, the per-position mismatch mask of all 20 positions is two LOP3s:
//     m = (A_w ^ A_g) | (B_w ^ B_g)
// and the mismatch count is one POPC -- no 2-bit fold, no shifts in the inner loop.
#include <algorithm>
#include <cstdlib>
#include <functional>
#include <vector>

#include "scan_common.cuh"
#include "join_blocks.cuh"

namespace guido {

constexpr int kSitesPerThread = 8;
constexpr int kSiteSlotBytes = 16;  // PAM-site table: one uint4 {position, hi plane | strand, lo plane, core key} per slot

// guide-side seed index (built per search call by the idx_* kernels below), structure of arrays:
// the search streams `keys` only and touches `guides` on a hit
struct OtIndex {
    const uint32_t *offsets;  // n_buckets + 1
    const uint32_t *keys;     // distal_hi | distal_lo << dl | core_mm << 2*dl
    const uint32_t *guides;   // guide index of the entry
    int cl, dl;               // core / distal length (cl + dl = 20)
    int field;                // bits per plane field of a key: dl + core_mm pseudo-positions
    int wide;                 // average bucket is large: walk buckets warp-wide
};

struct OtScanArgs {
    PlanesView pv;
    OtIndex ix;
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
    // PAM-site table (genome-side index): windows grouped by core key, buckets padded to 64 slots
    uint4 *tab_slots;                  // {global position, hi plane (+ strand in bit 31), lo plane, core key}; pads: x = ~0
    uint32_t *tab_counts, *tab_fill;   // per-bucket histogram -> starts (| pad count), fill cursors (build only)
    uint32_t tab_lo, tab_n;             // join: slots [tab_lo, tab_n) of the table
    uint32_t tab_key_lo, tab_key_span;  // table build: only core keys in [lo, lo + span) are tabled
};

// what the consumer of the front end does with a round of windows
constexpr int kModeScan = 0;      // brute force against every guide
constexpr int kModeIndex = 1;     // look up in the guide-side seed index
constexpr int kModeTabCount = 2;  // PAM-site table build, pass 1: histogram of core keys
constexpr int kModeTabFill = 3;   // PAM-site table build, pass 2: scatter into the buckets

// W bits starting at bit s of the 192-bit string w0:w1:w2 (s < 128)
__device__ __forceinline__ uint32_t extract_bits(uint64_t w0, uint64_t w1, uint64_t w2, int s, uint32_t wmask) {
    uint64_t a = (s < 64) ? w0 : w1, b = (s < 64) ? w1 : w2;
    int sh = s & 63;
    uint64_t v = a >> sh;
    if (sh) v |= b << (64 - sh);
    return (uint32_t)v & wmask;
}

// compare one round of sites held in the ring [head, head+n) against every guide
template <int RING>
__device__ __forceinline__ void compare_round(const OtScanArgs &a, const uint32_t *sA, const uint32_t *sB,
                                              const uint32_t *sQ, uint32_t head, uint32_t n) {
    uint32_t A[kSitesPerThread], B[kSitesPerThread];
    uint32_t valid = 0;
    const uint32_t cmp = a.op.cmp_mask;
#pragma unroll
    for (int k = 0; k < kSitesPerThread; k++) {
        uint32_t j = k * kThreads + threadIdx.x;
        uint32_t idx = (head + j) & (RING - 1);
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
                        uint32_t idx = (head + k * kThreads + threadIdx.x) & (RING - 1);
                        guido_ot_hit h;
                        h.guide = g; h.gpos = sQ[idx]; h.hi = sA[idx]; h.lo = sB[idx];
                        a.hits[slot] = h;
                    }
                }
            }
        }
    }
}

// per-warp scratch of the index lookup
struct LookupSmem {
    uint32_t start[kThreads / 32][32];
    uint32_t pref[kThreads / 32][33];
    uint32_t dist[kThreads / 32][32];
};

// Look the sites of the ring [head, head+n) up in the guide-side seed index.  A warp takes 32 sites
// at a time: every lane fetches one site's bucket range (32 independent loads), the ranges are
// flattened by a warp prefix sum, and the lanes then walk the flattened key list 4 x 32 keys per
// step (coalesced 4-byte loads inside a bucket, 4 independent loads per lane in flight), so lanes
// stay busy and HBM latency stays covered whatever the bucket sizes are.
// Hits of the index lookups are parked in shared memory and moved to the global list with ONE
// atomic per flush: a single global cursor serialises at L2 (~1 ns per atomic), which would cost
// ~10 ms for the 1.1e7 hits of the 100k-guide workload.
constexpr int kHitBuf = 384;
struct HitBuf {
    uint32_t n;
    unsigned long long base;
    unsigned long long visited;  // index entries compared by this CTA (statistics)
    guido_ot_hit rec[kHitBuf];
};

__device__ __forceinline__ void push_hit(const OtScanArgs &a, HitBuf &hb, const guido_ot_hit &h) {
    const uint32_t i = atomicAdd(&hb.n, 1u);
    if (i < kHitBuf) {
        hb.rec[i] = h;
    } else {  // parked list is full until the next flush: straight to the global list
        unsigned long long slot = atomicAdd(a.count, 1ull);
        if (slot < a.cap) a.hits[slot] = h;
    }
}

// whole CTA, after a barrier that ends the pushing phase
__device__ __forceinline__ void flush_hits(const OtScanArgs &a, HitBuf &hb, bool force) {
    const uint32_t n = min(hb.n, (uint32_t)kHitBuf);
    __syncthreads();  // everybody has looked at the fill level before anybody parks another hit
    if (n == 0 || (!force && n < kHitBuf / 2)) return;
    if (threadIdx.x == 0) hb.base = atomicAdd(a.count, (unsigned long long)n);
    __syncthreads();
    const unsigned long long base = hb.base;
    for (uint32_t i = threadIdx.x; i < n; i += kThreads)
        if (base + i < a.cap) a.hits[base + i] = hb.rec[i];
    __syncthreads();
    if (threadIdx.x == 0) hb.n = 0;
    __syncthreads();
}

// bucket of a core key: offsets[] holds starts that are multiples of 8 (the vector key loads stay
// aligned) with the amount of padding (0..7) in the low 3 bits
__device__ __forceinline__ void bucket_range(const OtIndex &ix, uint32_t key, uint32_t &start, uint32_t &len) {
    const uint32_t v = __ldg(&ix.offsets[key]), w = __ldg(&ix.offsets[key + 1]);
    start = v & ~7u;
    len = (w & ~7u) - start - (v & 7u);
}

template <int RING>
__device__ __forceinline__ void lookup_round(const OtScanArgs &a, const uint32_t *sA, const uint32_t *sB,
                                             const uint32_t *sQ, uint32_t head, uint32_t n, LookupSmem &ls,
                                             HitBuf &hb) {
    constexpr int U = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cl = a.ix.cl, dl = a.ix.dl;
    const uint32_t cmask = (1u << cl) - 1, dmask = (1u << dl) - 1;
    const int tot_max = a.op.total_max;
    const uint32_t *__restrict__ keys = a.ix.keys;
    uint32_t *w_start = ls.start[warp], *w_pref = ls.pref[warp], *w_dist = ls.dist[warp];
    const int F = a.ix.field;
    const uint32_t fmask = (1u << F) - 1;
    for (uint32_t b = warp * 32; b < n; b += kThreads) {
        const uint32_t j = b + lane;
        const uint32_t idx = (head + j) & (RING - 1);
        uint32_t start = 0, len = 0, dist = 0;
        if (j < n) {
            const uint32_t A = sA[idx], B = sB[idx];
            const uint32_t key = (((A >> dl) & cmask) << cl) | ((B >> dl) & cmask);
            bucket_range(a.ix, key, start, len);
            dist = (A & dmask) | ((B & dmask) << F);
        }
        uint32_t inc = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(kFull, inc, d);
            if (lane >= d) inc += t;
        }
        const uint32_t total = __shfl_sync(kFull, inc, 31);
        if (lane == 0) atomicAdd(&hb.visited, (unsigned long long)total);
        __syncwarp();
        w_start[lane] = start;
        w_pref[lane] = inc - len;
        w_dist[lane] = dist;
        if (lane == 31) w_pref[32] = total;
        __syncwarp();
        uint32_t cur = 0;
        for (uint32_t t0 = lane; t0 < total; t0 += 32 * U) {
            uint32_t pos[U], site[U], kv[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const uint32_t t = t0 + 32 * u;
                if (t < total) {
                    while (t >= w_pref[cur + 1]) cur++;
                    site[u] = cur;
                    pos[u] = w_start[cur] + (t - w_pref[cur]);
                } else {
                    site[u] = 0xffffffffu;
                    pos[u] = 0;
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) kv[u] = (site[u] != 0xffffffffu) ? __ldg(&keys[pos[u]]) : 0u;
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (site[u] == 0xffffffffu) continue;
                const uint32_t x = kv[u] ^ w_dist[site[u]];
                const uint32_t m = (x | (x >> F)) & fmask;
                if ((int)__popc(m) <= tot_max) {
                    const uint32_t si = (head + b + site[u]) & (RING - 1);
                    guido_ot_hit h;
                    h.guide = __ldg(&a.ix.guides[pos[u]]); h.gpos = sQ[si]; h.hi = sA[si]; h.lo = sB[si];
                    push_hit(a, hb, h);
                }
            }
        }
    }
}

// Same lookup for LARGE buckets (>= ~12 guides per bucket, e.g. 100k guides): the warp walks one
// bucket at a time with all 32 lanes -- coalesced 128-byte key loads and no per-key search for
// the owning site -- and keeps the loads of 8 buckets in flight together.
template <int RING>
__device__ __forceinline__ void lookup_round_wide(const OtScanArgs &a, const uint32_t *sA, const uint32_t *sB,
                                                  const uint32_t *sQ, uint32_t head, uint32_t n, HitBuf &hb) {
    constexpr int U = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cl = a.ix.cl, dl = a.ix.dl, F = a.ix.field;
    const uint32_t cmask = (1u << cl) - 1, dmask = (1u << dl) - 1, fmask = (1u << F) - 1;
    const int tot_max = a.op.total_max;
    const uint32_t *__restrict__ keys = a.ix.keys;
    for (uint32_t b = warp * 32; b < n; b += kThreads) {
        const uint32_t j = b + lane;
        uint32_t start = 0, len = 0, dist = 0;
        if (j < n) {
            const uint32_t idx = (head + j) & (RING - 1);
            const uint32_t A = sA[idx], B = sB[idx];
            const uint32_t key = (((A >> dl) & cmask) << cl) | ((B >> dl) & cmask);
            bucket_range(a.ix, key, start, len);
            dist = (A & dmask) | ((B & dmask) << F);
        }
        {
            const uint32_t warp_total = __reduce_add_sync(kFull, len);
            if (lane == 0) atomicAdd(&hb.visited, (unsigned long long)warp_total);
        }
        for (int s0 = 0; s0 < 32; s0 += U) {
            uint32_t st[U], ln[U], ds[U], maxlen = 0;
#pragma unroll
            for (int u = 0; u < U; u++) {
                st[u] = __shfl_sync(kFull, start, s0 + u);
                ln[u] = __shfl_sync(kFull, len, s0 + u);
                ds[u] = __shfl_sync(kFull, dist, s0 + u);
                maxlen = max(maxlen, ln[u]);
            }
            // a lane takes keys off, off+1 of each bucket: one aligned 8-byte load, 64 keys per warp
            // step.  The first step is straight-line code (a bucket rarely holds more than 64 keys).
            auto step = [&](const uint32_t off) {
                uint2 kv[U];
#pragma unroll
                for (int u = 0; u < U; u++)
                    kv[u] = (off < ln[u]) ? __ldg(reinterpret_cast<const uint2 *>(keys + st[u] + off)) : make_uint2(0u, 0u);
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const uint32_t x0 = kv[u].x ^ ds[u], x1 = kv[u].y ^ ds[u];
                    const bool h0 = off < ln[u] && (int)__popc((x0 | (x0 >> F)) & fmask) <= tot_max;
                    const bool h1 = off + 1 < ln[u] && (int)__popc((x1 | (x1 >> F)) & fmask) <= tot_max;
                    if (h0 | h1) {
                        const uint32_t si = (head + b + s0 + u) & (RING - 1);
                        guido_ot_hit h;
                        h.gpos = sQ[si]; h.hi = sA[si]; h.lo = sB[si];
                        if (h0) { h.guide = __ldg(&a.ix.guides[st[u] + off]); push_hit(a, hb, h); }
                        if (h1) { h.guide = __ldg(&a.ix.guides[st[u] + off + 1]); push_hit(a, hb, h); }
                    }
                }
            };
            step(2 * lane);
            for (uint32_t off = 2 * lane + 64; off < maxlen; off += 64) step(off);
        }
    }
}

// one window's walk through its bucket: current 8 keys in registers, the next 8 in flight
struct BucketWalk {
    const uint32_t *p;   // bucket start (32-byte aligned)
    uint32_t len, off;   // true length, keys consumed so far
    uint32_t dist, idx;  // the window's distal planes in key layout; its ring slot
    unsigned long long n0, n1, n2, n3;
};

// kReuse: neighbouring windows walk the same bucket (join over the sorted site table), so the
// sector is worth keeping in L1; the streaming lookup reads every sector once
template <bool kReuse>
__device__ __forceinline__ void walk_load(BucketWalk &w, uint32_t off) {
    if constexpr (kReuse)
        asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(w.n0), "=l"(w.n1), "=l"(w.n2), "=l"(w.n3) : "l"(w.p + off));
    else
        asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(w.n0), "=l"(w.n1), "=l"(w.n2), "=l"(w.n3) : "l"(w.p + off));
}

template <int RING>
__device__ __forceinline__ void walk_init(const OtScanArgs &a, const uint32_t *sA, const uint32_t *sB, uint32_t head,
                                          uint32_t j, uint32_t n, BucketWalk &w, unsigned long long &visited) {
    const int cl = a.ix.cl, dl = a.ix.dl, F = a.ix.field;
    const uint32_t cmask = (1u << cl) - 1, dmask = (1u << dl) - 1;
    w.len = 0; w.off = 0; w.p = a.ix.keys; w.dist = 0; w.idx = 0;
    w.n0 = w.n1 = w.n2 = w.n3 = 0;
    if (j >= n) return;
    w.idx = (head + j) & (RING - 1);
    const uint32_t A = sA[w.idx], B = sB[w.idx];
    const uint32_t key = (((A >> dl) & cmask) << cl) | ((B >> dl) & cmask);
    uint32_t start;
    bucket_range(a.ix, key, start, w.len);
    w.dist = (A & dmask) | ((B & dmask) << F);
    w.p = a.ix.keys + start;
    visited += w.len;
    if (w.len) walk_load<false>(w, 0);
    // pull the rest of the bucket towards L2 right away (no registers tied up): the later sectors
    // of the walk then cost an L2 hit instead of a DRAM round trip
    for (uint32_t o = 32; o < w.len; o += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(w.p + o));
}

// compare the 8 keys in registers (the next 8 are requested first); hits are parked
// rec(idx) -> {position, hi plane, lo plane} of the window a hit belongs to
template <bool kReuse, typename Rec>
__device__ __forceinline__ void walk_step(const OtScanArgs &a, const Rec &rec, BucketWalk &w, HitBuf &hb) {
    const int F = a.ix.field;
    const uint32_t fmask = (1u << F) - 1;
    const int tot_max = a.op.total_max;
    uint32_t kv[8];
    kv[0] = (uint32_t)w.n0; kv[1] = (uint32_t)(w.n0 >> 32); kv[2] = (uint32_t)w.n1; kv[3] = (uint32_t)(w.n1 >> 32);
    kv[4] = (uint32_t)w.n2; kv[5] = (uint32_t)(w.n2 >> 32); kv[6] = (uint32_t)w.n3; kv[7] = (uint32_t)(w.n3 >> 32);
    const uint32_t off = w.off;
    if (off + 8 < w.len) walk_load<kReuse>(w, off + 8);
    w.off = off + 8;
    bool any = false;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t x = kv[k] ^ w.dist;
        any |= (int)__popc((x | (x >> F)) & fmask) <= tot_max;
    }
    if (any) {  // rare (6e-4 per key): redo with the validity check and park the hits
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t x = kv[k] ^ w.dist;
            if (off + k < w.len && (int)__popc((x | (x >> F)) & fmask) <= tot_max) {
                guido_ot_hit h;
                h.guide = __ldg(&a.ix.guides[(w.p - a.ix.keys) + off + k]);
                const uint3 win = rec(w.idx);
                h.gpos = win.x; h.hi = win.y; h.lo = win.z;
                push_hit(a, hb, h);
            }
        }
    }
}

// Default lookup: ONE THREAD PER WINDOW.  A lane walks its own bucket with 256-bit loads (8 keys,
// one full 32-byte sector per load; buckets start on 32-byte boundaries), so a warp instruction
// advances 32 windows at once and the per-window overhead of the warp-per-window variants
// (shuffles, per-key validity, loop control) disappears: ~5 issue slots per key compared.  The
// kernel is latency-bound on the key loads (profiles/r1_offtarget.md), so the next sector of the
// bucket is requested before the current one is compared.  (Walking two windows per thread at once
// was measured slower: 21.6 vs 19.6 ms, register spills and extra divergence.)  The padding keys
// at a bucket's end are not filtered in the hot loop -- only in the rare hit path.
template <int RING>
__device__ __forceinline__ void lookup_round_thread(const OtScanArgs &a, const uint32_t *sA, const uint32_t *sB,
                                                    const uint32_t *sQ, uint32_t head, uint32_t n, HitBuf &hb) {
    unsigned long long visited = 0;
    for (uint32_t j = threadIdx.x; j < n; j += kThreads) {
        BucketWalk w;
        walk_init<RING>(a, sA, sB, head, j, n, w, visited);
        const auto rec = [&](uint32_t idx) { return make_uint3(sQ[idx], sA[idx], sB[idx]); };
        while (w.off < w.len) walk_step<false>(a, rec, w, hb);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) visited += __shfl_xor_sync(kFull, visited, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&hb.visited, visited);
}

// bit j of x (j < 16) -> bit 2j
__device__ __forceinline__ uint32_t spread_bits(uint32_t x) {
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    return (x | (x << 1)) & 0x55555555u;
}

// PAM-site table build: the core key of every window of the round is counted (pass 1) or the
// window is scattered to its bucket (pass 2).  Order inside a bucket is arbitrary; hit lists are
// sorted canonically afterwards.
template <int RING, bool kFill>
__device__ __forceinline__ void table_round(const OtScanArgs &a, const uint32_t *sA, const uint32_t *sB,
                                            const uint32_t *sQ, uint32_t head, uint32_t n) {
    const int cl = a.ix.cl, dl = a.ix.dl;
    const uint32_t cmask = (1u << cl) - 1;
    uint32_t tabled = 0;
    for (uint32_t j = threadIdx.x; j < n; j += kThreads) {
        const uint32_t idx = (head + j) & (RING - 1);
        const uint32_t A = sA[idx], B = sB[idx];
        const uint32_t key = (((A >> dl) & cmask) << cl) | ((B >> dl) & cmask);
        if (key - a.tab_key_lo >= a.tab_key_span) continue;  // another rank's share of the key space
        if constexpr (!kFill) {
            atomicAdd(&a.tab_counts[key], 1u);
            tabled++;
        } else {
            const uint32_t slot = (a.tab_counts[key] & ~63u) + atomicAdd(&a.tab_fill[key], 1u);
            a.tab_slots[slot] = make_uint4(sQ[idx], A, B, key);
        }
    }
    if constexpr (!kFill) {  // [1]: windows that fall into this handle's share of the key space
        tabled = __reduce_add_sync(kFull, tabled);
        if ((threadIdx.x & 31) == 0 && tabled && a.site_count) atomicAdd(a.site_count + 1, (unsigned long long)tabled);
    }
}

template <int kMode>
struct OtCfg {
    // the brute-force consumer wants 8 sites per thread in registers (2048 per round); the other
    // consumers only need dense warps, so a smaller ring buys occupancy (more loads in flight)
    static constexpr int kRing = kMode == kModeScan ? 4096 : 2048;
    static constexpr int kRound = kMode == kModeScan ? kThreads * kSitesPerThread : 1024;
};

template <int kMode>
__global__ void __launch_bounds__(kThreads, kMode == kModeScan ? 3 : 5) ot_kernel(const OtScanArgs a) {
    constexpr bool kIndex = kMode == kModeIndex;
    constexpr int RING = OtCfg<kMode>::kRing, ROUND = OtCfg<kMode>::kRound;
    extern __shared__ uint32_t s_dyn[];
    uint32_t *sA = s_dyn, *sB = s_dyn + RING, *sQ = s_dyn + 2 * RING;
    __shared__ LookupSmem s_lookup;  // the index consumer's scratch (dead in the brute-force instantiation)
    __shared__ HitBuf s_hits;
    __shared__ uint32_t s_warp[kThreads / 32];
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { s_hits.n = 0; s_hits.visited = 0; }
    __syncthreads();
    const int P = a.op.pam.len;
    const int W = kProto + P;
    const uint32_t wmask = (1u << W) - 1;
    const int raw = a.op.raw_mode;
    uint32_t head = 0, qlen = 0;     // ring state (uniform across the CTA)
    unsigned long long n_sites = 0;

    for (uint32_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int64_t tile_block = a.first_block + (int64_t)tile * kThreads;
        const uint32_t t0 = (uint32_t)(tile_block * 64), t1 = t0 + kTileBases;
        int r_lo, r_hi;
        bool skip;
        tile_runs(a.pv, t0, t1, r_lo, r_hi, skip);
        if (skip) continue;
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

        // append this tile's sites to the ring; run full rounds whenever enough are queued
        while (true) {
            uint32_t total;
            const uint32_t mine = __popcll(hf) + __popcll(hr);
            uint32_t off = block_exclusive_scan(mine, s_warp, &total);
            if (total == 0) break;
            const uint32_t space = RING - qlen;
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
                    uint32_t idx = (head + qlen + off) & (RING - 1);
                    sA[idx] = hi_bits; sB[idx] = lo_bits; sQ[idx] = q;
                    off++;
                }
                if (strand == 0) hf = m; else hr = m;
            }
            qlen += appended;
            n_sites += (threadIdx.x == 0) ? appended : 0;
            __syncthreads();
            while (qlen >= ROUND) {
                if constexpr (kIndex) {
                    if (a.ix.wide == 2) lookup_round_thread<RING>(a, sA, sB, sQ, head, ROUND, s_hits);
                    else if (a.ix.wide == 1) lookup_round_wide<RING>(a, sA, sB, sQ, head, ROUND, s_hits);
                    else lookup_round<RING>(a, sA, sB, sQ, head, ROUND, s_lookup, s_hits);
                    __syncthreads();
                    flush_hits(a, s_hits, false);
                } else if constexpr (kMode == kModeScan) {
                    compare_round<RING>(a, sA, sB, sQ, head, ROUND);
                } else {
                    table_round<RING, kMode == kModeTabFill>(a, sA, sB, sQ, head, ROUND);
                }
                head = (head + ROUND) & (RING - 1);
                qlen -= ROUND;
            }
            __syncthreads();
            if (appended == total) break;
        }
    }
    __syncthreads();
    if constexpr (kIndex) {
        if (qlen) {
            if (a.ix.wide == 2) lookup_round_thread<RING>(a, sA, sB, sQ, head, qlen, s_hits);
            else if (a.ix.wide == 1) lookup_round_wide<RING>(a, sA, sB, sQ, head, qlen, s_hits);
            else lookup_round<RING>(a, sA, sB, sQ, head, qlen, s_lookup, s_hits);
        }
        __syncthreads();
        flush_hits(a, s_hits, true);
    } else if constexpr (kMode == kModeScan) {
        if (qlen) compare_round<RING>(a, sA, sB, sQ, head, qlen);
    } else {
        if (qlen) table_round<RING, kMode == kModeTabFill>(a, sA, sB, sQ, head, qlen);
    }
    if (threadIdx.x == 0 && a.site_count) {  // [0] windows examined, [1] index entries compared
        atomicAdd(a.site_count, n_sites);
        if constexpr (kIndex) atomicAdd(a.site_count + 1, s_hits.visited);
    }
}

constexpr uint32_t kJoinPerThread = 4;  // windows per thread between two CTA barriers

template <int kMode>
static int launch_ot_kernel(DeviceState *ds, OtScanArgs &a, cudaStream_t stream) {
    a.first_block = a.r.lo / 64;
    int64_t last_block = ((int64_t)a.r.hi - 1) / 64;
    a.n_tiles = (uint32_t)((last_block - a.first_block + 1 + kThreads - 1) / kThreads);
    const int smem = 3 * OtCfg<kMode>::kRing * (int)sizeof(uint32_t);
    GUIDO_CUDA(cudaFuncSetAttribute(ot_kernel<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = 0;
    GUIDO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ot_kernel<kMode>, kThreads, smem));
    int grid = ds->sm_count * std::max(per_sm, 1);
    if ((uint32_t)grid > a.n_tiles) grid = (int)a.n_tiles;
    constexpr bool kTimed = kMode == kModeScan || kMode == kModeIndex;
    if (kTimed) kev_begin(ds, stream);
    ot_kernel<kMode><<<grid, kThreads, smem, stream>>>(a);
    if (kTimed) kev_end(ds, stream);
    GUIDO_CUDA(cudaGetLastError());
    return GUIDO_OK;
}

// ------------------------------------------------------------------------- join over the site table
//
// With the PAM-valid windows of the genome grouped by core key (the PAM-site table) and the guides
// filed under the same keys (the seed index below), the search is a bucket-by-bucket join.  One
// thread per window, windows in table order: the 32 lanes of a warp sit in the same bucket (or in
// neighbouring ones), so the 256-bit key loads of a warp hit the same sectors -- the index is read
// from HBM once (L1/L2 serve the rest) instead of once per window, and what remains is the
// compare itself: XOR, fold, POPC, compare per key.
__device__ __forceinline__ void join_init(const OtScanArgs &a, uint32_t i, BucketWalk &w, unsigned long long &visited) {
    const int cl = a.ix.cl, dl = a.ix.dl, F = a.ix.field;
    const uint32_t cmask = (1u << cl) - 1, dmask = (1u << dl) - 1;
    w.len = 0; w.off = 0; w.p = a.ix.keys; w.dist = 0; w.idx = i;
    w.n0 = w.n1 = w.n2 = w.n3 = 0;
    if (i >= a.tab_n) return;
    const uint4 slot = __ldg(&a.tab_slots[i]);
    if (slot.x == 0xffffffffu) return;  // pad slot at the end of a bucket
    const uint32_t A = slot.y, B = slot.z;
    const uint32_t key = (((A >> dl) & cmask) << cl) | ((B >> dl) & cmask);
    uint32_t start;
    bucket_range(a.ix, key, start, w.len);
    w.dist = (A & dmask) | ((B & dmask) << F);
    w.p = a.ix.keys + start;
    visited += w.len;
    if (w.len) walk_load<true>(w, 0);
}


__global__ void __launch_bounds__(kThreads, 5) ot_join_kernel(const OtScanArgs a) {
    __shared__ HitBuf s_hits;
    if (threadIdx.x == 0) { s_hits.n = 0; s_hits.visited = 0; }
    __syncthreads();
    // a CTA owns a contiguous stretch of the table, so consecutive chunks reuse the bucket in L1
    constexpr uint32_t kChunk = kJoinPerThread * kThreads;
    const uint32_t n_chunks = (a.tab_n - a.tab_lo + kChunk - 1) / kChunk;
    const uint32_t per_cta = (n_chunks + gridDim.x - 1) / gridDim.x;
    const uint32_t c0 = min(blockIdx.x * per_cta, n_chunks), c1 = min(c0 + per_cta, n_chunks);
    unsigned long long visited = 0;
    for (uint32_t c = c0; c < c1; c++) {
        for (uint32_t k = 0; k < kJoinPerThread; k++) {
            BucketWalk w;
            join_init(a, a.tab_lo + c * kChunk + k * kThreads + threadIdx.x, w, visited);
            const auto rec = [&](uint32_t idx) { const uint4 t = __ldg(&a.tab_slots[idx]); return make_uint3(t.x, t.y, t.z); };
            while (w.off < w.len) walk_step<true>(a, rec, w, s_hits);
        }
        __syncthreads();
        flush_hits(a, s_hits, false);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) visited += __shfl_xor_sync(kFull, visited, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_hits.visited, visited);
    __syncthreads();
    flush_hits(a, s_hits, true);
    if (threadIdx.x == 0 && a.site_count) {
        atomicAdd(a.site_count + 1, s_hits.visited);
    }
}

int launch_ot_scan(DeviceState *ds, const PlanesView &pv, const OtParams &op, const ScanRange &r,
                   const uint2 *d_guides, int64_t n_guides, guido_ot_hit *d_hits, uint64_t cap,
                   unsigned long long *d_count, uint8_t *d_flags, unsigned long long *d_site_count,
                   cudaStream_t stream, int *n_launches) {
    if (r.hi <= r.lo || n_guides == 0) return GUIDO_OK;
    OtScanArgs a;
    a.pv = pv; a.op = op; a.r = r;
    a.ix = OtIndex{nullptr, nullptr, nullptr, 0, 0, 0, 0};
    a.guides = d_guides; a.n_guides = (uint32_t)n_guides;
    a.hits = d_hits; a.cap = cap; a.count = d_count; a.flags = d_flags; a.site_count = d_site_count;
    a.tab_slots = nullptr; a.tab_counts = a.tab_fill = nullptr; a.tab_n = 0;
    int rc = launch_ot_kernel<kModeScan>(ds, a, stream);
    if (rc) return rc;
    if (n_launches) *n_launches += 1;
    return GUIDO_OK;
}

// ------------------------------------------------------------------------- seed index build
//
// Key of a guide / window = its core_len PAM-proximal bases as (hi plane << cl) | lo plane.
// Every guide is filed under ALL keys within <= core_mm substitutions of its own core (436 keys
// for core 10 / 2 mismatches), so a window finds, under its own key, exactly the guides whose
// core is within the budget -- the bowtie seed rule -- and only the distal bases are left to
// compare.  A substitution XORs the 2-bit code with 1..3, so the variants are the (xh, xl) plane
// masks with popc(xh | xl) <= core_mm; the host enumerates them once per call.

struct IdxArgs {
    const uint2 *guides;
    uint32_t n_guides;
    const uint32_t *variants;  // xh | xl << 16
    uint32_t n_variants;
    int cl, dl, field;
    uint32_t *counts;          // n_buckets + 1 (pass 1), becomes offsets after the scan
    uint32_t *fill;            // n_buckets
    uint32_t *keys, *ids;      // n_entries each
    uint32_t slice_bits, slice;  // fill pass: only buckets whose top slice_bits key bits == slice
    uint16_t group_off[33];      // variants sorted by the top slice_bits of their hi-plane mask
    uint32_t part_shift, part_ix; // histogram pass: only keys with key >> part_shift == part_ix (when part_on)
    bool part_on;
    uint32_t pad_mask;         // bucket starts are multiples of pad_mask + 1 (7, or 0: no padding)
    int entries;               // 1: one {letters | bias, guide} entry per key, the format of the block join (join_blocks.cu)
    uint32_t cbias;            // entries: 7 - total_max, added to the variant's core mismatch count
    uint32_t core_max;         // entries: variants with this many core mismatches go to the front of a bucket
    uint32_t *back;            // entries: n_buckets cursors that start at the bucket ends
};

// One warp per guide: the guide is read once, the lanes walk the variant list (staged in shared
// memory) -- no per-entry division, and the scatter is the only memory traffic that matters.
constexpr int kMaxVariants = 4096;  // core 12 / 3 mismatches: 6535 would not fit -> falls back to global reads

template <bool kFill>
__global__ void __launch_bounds__(256) idx_build_kernel(const IdxArgs a) {
    __shared__ uint32_t s_var[kMaxVariants];
    const bool staged = a.n_variants <= (uint32_t)kMaxVariants;
    if (staged)
        for (uint32_t v = threadIdx.x; v < a.n_variants; v += 256) s_var[v] = __ldg(&a.variants[v]);
    __syncthreads();
    const uint32_t cmask = (1u << a.cl) - 1, dmask = (1u << a.dl) - 1;
    const uint32_t lane = threadIdx.x & 31, warps = gridDim.x * 8u;
    uint2 *ent = reinterpret_cast<uint2 *>(a.keys);  // entry format: {letters | bias, guide} per key
    for (uint32_t g = blockIdx.x * 8u + (threadIdx.x >> 5); g < a.n_guides; g += warps) {
        const uint2 G = __ldg(&a.guides[g]);
        const uint32_t ch = (G.x >> a.dl) & cmask, cl_ = (G.y >> a.dl) & cmask;
        const uint32_t dh = G.x & dmask, dlo = G.y & dmask;
        // entry format: the distal bases as 2-bit letter codes, base j at bits 2j (= 2 * hi + lo)
        const uint32_t letters = kFill && a.entries ? (spread_bits(dlo) | (spread_bits(dh) << 1)) : 0u;
        // A fill pass writes the buckets whose top `slice_bits` key bits equal `slice`; the variants
        // are grouped by the same bits of their hi-plane mask, so the group (top bits of the
        // guide's core ^ slice) holds exactly this pass's variants: every lane of every iteration
        // does a store, and the eight passes together walk the variant list once.
        uint32_t v0 = 0, v1 = a.n_variants;
        if constexpr (kFill) {
            const uint32_t grp = ((ch >> (a.cl - a.slice_bits)) ^ a.slice) & ((1u << a.slice_bits) - 1u);
            v0 = a.group_off[grp]; v1 = a.group_off[grp + 1];
        }
        for (uint32_t v = v0 + lane; v < v1; v += 32) {
            const uint32_t var = staged ? s_var[v] : __ldg(&a.variants[v]);
            const uint32_t xh = var & 0xffff, xl = var >> 16;
            const uint32_t key = ((ch ^ xh) << a.cl) | (cl_ ^ xl);
            if constexpr (!kFill) {
                if (!a.part_on || (key >> a.part_shift) == a.part_ix) atomicAdd(&a.counts[key], 1u);
            } else {
                // fill[] starts out as the bucket starts (scan_fixup_kernel): the atomic IS the slot.
                // Entry format: the variants with the most core mismatches (the bulk: 405 of 436 for
                // 10 / 2) fill a bucket from the front, the others from the back, so whole passes of
                // the block join see one bias and may take the short compare tree.
                const uint32_t cmm = (uint32_t)__popc(xh | xl);
                const uint32_t pos = a.entries && cmm != a.core_max ? atomicSub(&a.back[key], 1u) - 1u : atomicAdd(&a.fill[key], 1u);
                if (a.entries) {
                    // the guide's distal letters and the biased core mismatch count of this variant,
                    // c' = c + 7 - total_max (a hit is d + c' <= 7); one 8-byte store per entry
                    ent[pos] = make_uint2(letters | (min(cmm + a.cbias, 15u) << 20), g);
                } else {
                    // key = distal planes in two fields of `field` bits; the core mismatches of this
                    // variant are appended to the hi field as that many 1-bits, which the window's
                    // zeros turn into mismatches -- popc of the folded XOR is the TOTAL mismatch count
                    a.keys[pos] = dh | (((1u << cmm) - 1u) << a.dl) | (dlo << a.field);
                    a.ids[pos] = g;
                }
            }
        }
    }
}

// ---- bucket sizes without 436 atomics per guide ---------------------------------------------------
//
// counts[key] = number of guides whose core is within <= core_mm substitutions of `key` = the exact
// histogram of the guides' own cores convolved with a Hamming ball.  The histogram pass above does
// that with one atomic per (guide, variant) -- 43.6 M L2 atomics, 0.19 ms that nothing else in the
// search can overlap.  The convolution factorises over the core positions:
//     g_j[key] = number of guides that differ from key in exactly j of the positions handled so far
//     position p:  g_j[key] += sum over the 3 other letters c of g_{j-1}[key with position p := c]
// (levels updated from the highest down, so g_{j-1} is still the previous step's).  Two kernels, the
// low half of the positions and the high half: a CTA holds every key that differs from its tile only
// in "its" positions -- 4^5 keys x (core_mm + 1) levels for a 10-base core -- in shared memory and
// runs its positions there.  One atomic per guide (the exact histogram), ~40 MB of L2 traffic.
constexpr int kDpMaxKeys = 1024;  // 4^5: cores up to 10 bases

__global__ void __launch_bounds__(256) core_hist_kernel(const uint2 *guides, uint32_t n_guides, int cl, int dl, uint32_t *level0) {
    const uint32_t g = blockIdx.x * 256u + threadIdx.x;
    if (g >= n_guides) return;
    const uint2 G = __ldg(&guides[g]);
    const uint32_t cmask = (1u << cl) - 1;
    atomicAdd(&level0[(((G.x >> dl) & cmask) << cl) | ((G.y >> dl) & cmask)], 1u);
}

// kHigh = false: positions [0, h) of the core, one tile per value of the other positions' bits;
// levels[0] is read, levels[1 .. cm] are written.  kHigh = true: positions [h, cl); all levels are
// read and their sum goes to counts[] (0 outside the handle's share of the key space).
template <bool kHigh>
__global__ void __launch_bounds__(256) ball_count_kernel(uint32_t *levels, uint32_t n_buckets, int cl, int h, int cm,
                                                         uint32_t *counts, bool part_on, uint32_t part_shift, uint32_t part_ix) {
    __shared__ uint32_t s[4][kDpMaxKeys];
    const int w = kHigh ? cl - h : h;       // positions handled here
    const int o = kHigh ? h : cl - h;       // positions fixed by the tile
    const uint32_t n_local = 1u << (2 * w);
    const uint32_t t_hi = blockIdx.x >> o, t_lo = blockIdx.x & ((1u << o) - 1u);
    auto key_of = [&](uint32_t x) {
        const uint32_t x_hi = x >> w, x_lo = x & ((1u << w) - 1u);
        const uint32_t hi = kHigh ? ((x_hi << h) | t_hi) : ((t_hi << h) | x_hi);
        const uint32_t lo = kHigh ? ((x_lo << h) | t_lo) : ((t_lo << h) | x_lo);
        return (hi << cl) | lo;
    };
    for (uint32_t x = threadIdx.x; x < n_local; x += 256) {
        const uint32_t key = key_of(x);
        s[0][x] = levels[key];
        for (int j = 1; j <= cm; j++) s[j][x] = kHigh ? levels[(size_t)j * n_buckets + key] : 0u;
    }
    for (int p = 0; p < w; p++) {
        const uint32_t b_lo = 1u << p, b_hi = 1u << (w + p);
        for (int j = cm; j >= 1; j--) {
            __syncthreads();  // level j - 1 is complete for this step (and the loads above are done)
            for (uint32_t x = threadIdx.x; x < n_local; x += 256)
                s[j][x] += s[j - 1][x ^ b_lo] + s[j - 1][x ^ b_hi] + s[j - 1][x ^ b_lo ^ b_hi];
        }
    }
    __syncthreads();
    for (uint32_t x = threadIdx.x; x < n_local; x += 256) {
        const uint32_t key = key_of(x);
        if (kHigh) {
            uint32_t c = 0;
            for (int j = 0; j <= cm; j++) c += s[j][x];
            counts[key] = (!part_on || (key >> part_shift) == part_ix) ? c : 0u;
        } else {
            for (int j = 1; j <= cm; j++) levels[(size_t)j * n_buckets + key] = s[j][x];
        }
    }
}

// in-place exclusive scan of n u32 values: 4096 per block, then the block sums, then the fix-up
// With pad = 8 or 32 every value is rounded up to a multiple of pad before it is scanned (bucket
// starts stay aligned for the 256-bit key loads / for whole 32-key blocks) and the padding
// (0..pad-1) is kept in the low bits of the result: start = v & ~(pad-1), true length = next
// start - start - (v & (pad-1)).
__global__ void __launch_bounds__(1024) scan_blocks_kernel(uint32_t *data, uint32_t n, uint32_t *block_sums,
                                                            uint32_t pad) {
    __shared__ uint32_t s_w[32];
    const uint32_t base = blockIdx.x * 4096u + threadIdx.x * 4u;
    uint32_t v[4], odd[4], sum = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t c = (base + i < n) ? data[base + i] : 0;
        odd[i] = pad ? ((pad - (c & (pad - 1u))) & (pad - 1u)) : 0u;
        v[i] = c + odd[i];
        sum += v[i];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(kFull, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_w[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(kFull, wi, d); if (lane >= d) wi += t; }
        s_w[lane] = wi - w;
        if (lane == 31) block_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    uint32_t run = s_w[warp] + inc - sum;
#pragma unroll
    for (int i = 0; i < 4; i++) { if (base + i < n) data[base + i] = run | odd[i]; run += v[i]; }
}

__global__ void __launch_bounds__(1024) scan_sums_kernel(uint32_t *block_sums, uint32_t n_blocks) {
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n_blocks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_blocks ? block_sums[i] : 0;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(kFull, inc, d); if (lane >= d) inc += t; }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_w[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(kFull, wi, d); if (lane >= d) wi += t; }
            s_w[lane] = wi - w;
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        if (i < n_blocks) block_sums[i] = carry + s_w[warp] + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_w[31] + inc;
        __syncthreads();
    }
}

// `cursor` (optional, n - 1 entries): the bucket starts without the padding bits -- the fill pass
// takes slots from it with ONE returning atomic per entry instead of a load plus an atomic
__global__ void __launch_bounds__(1024) scan_fixup_kernel(uint32_t *data, uint32_t n, const uint32_t *block_sums,
                                                           uint32_t *cursor, uint32_t pad_mask) {
    const uint32_t base = blockIdx.x * 4096u + threadIdx.x * 4u;
    const uint32_t add = block_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (base + i < n) {
            const uint32_t v = data[base + i] + add;
            data[base + i] = v;
            if (cursor && base + i + 1 < n) cursor[base + i] = v & ~pad_mask;
        }
}

// in-place exclusive scan of data[0 .. n) on `stream` (sums: scratch of ceil(n / 4096) words; cursor,
// optional, n - 1 entries: a second copy of the result for fill cursors) -- used by sort.cu as well
int exclusive_scan_u32(uint32_t *data, uint32_t n, uint32_t *sums, uint32_t *cursor, cudaStream_t stream) {
    if (n == 0) return GUIDO_OK;
    const uint32_t n_blocks = (n + 4095) / 4096;
    scan_blocks_kernel<<<n_blocks, 1024, 0, stream>>>(data, n, sums, 0);
    scan_sums_kernel<<<1, 1024, 0, stream>>>(sums, n_blocks);
    scan_fixup_kernel<<<n_blocks, 1024, 0, stream>>>(data, n, sums, cursor, 0);
    GUIDO_CUDA(cudaGetLastError());
    return GUIDO_OK;
}

// back[key] = end of bucket key's entries (offsets carry the pad count of a bucket in their low bits)
__global__ void __launch_bounds__(256) idx_back_cursor_kernel(const uint32_t *offsets, uint32_t n_buckets, uint32_t pad_mask,
                                                               uint32_t *back) {
    const uint32_t k = blockIdx.x * 256u + threadIdx.x;
    if (k < n_buckets) back[k] = (offsets[k + 1] & ~pad_mask) - (offsets[k] & pad_mask);
}

// ---- entries by GATHER: bucket-major, no atomics, coalesced stores ---------------------------------
//
// The scatter above costs one returning L2 atomic and one isolated 8-byte store per entry (43.6 M of
// each for 100 k guides) and has to walk the guides once per L2-sized slice.  The same index can be
// read off the guides' OWN cores: bucket b holds the guides whose core is one of the 436 neighbours
// of b (<= core_mm substitutions), only ~9 % of all cores carry a guide, and which ones do is a
// 1 Mbit map.  So: guides sorted by core (counting sort on the exact histogram the Hamming-ball
// convolution starts from), a presence bitmap, and per bucket a warp that probes the 436 neighbour
// bits, compacts the occupied ones IN VARIANT ORDER (largest core-mismatch class first, which is the
// order the block join wants) and copies their guides into the bucket: consecutive lanes write
// consecutive entries.
//
// The core positions are split in halves A (0..4) and B (5..9).  A CTA owns buckets that share their
// B half, so all its probes fall into the 106 rows "B half within <= 2 substitutions" of the bitmap;
// a row is 1024 bits (the A half: word = hi plane bits, bit = lo plane bits), 13.6 KB for all rows,
// staged in shared memory once per CTA.  Lane l owns variants 14 l .. 14 l + 13 for the whole
// kernel: row slot and A-half XOR masks sit in 14 registers, a probe is XOR, shift, LDS, shift.
constexpr int kGatherPerLane = 14;                       // 32 x 14 = 448 >= 436 variants (core 10 / <= 2)
constexpr int kGatherMaxVar = 32 * kGatherPerLane;
constexpr int kGatherMaxRows = 106;                      // B halves within <= 2 substitutions; + 1 all-zero row
constexpr int kGatherWarps = 8;
constexpr int kGatherSub = 4;                            // CTAs per B tile: 256 buckets each

struct GatherArgs {
    const uint32_t *present;     // bit `key` set: some guide has this core (key = hi plane << 10 | lo plane)
    const uint2 *core_rec;       // per occupied core: {distal letters, guide} of the guide that stands for it
    const uint32_t *ent_off;     // n_buckets + 1: first entry of each bucket (even) | pad flag
    uint2 *ent;
    const uint32_t *probe;       // kGatherMaxVar words: row slot << 10 | A-half hi mask << 5 | A-half lo mask
    const uint32_t *var_mask;    // kGatherMaxVar words: key XOR mask (20 bits) | bias << 20
    const uint32_t *row_mask;    // kGatherMaxRows words: B-half hi mask << 5 | lo mask of each row
    uint32_t n_rows;
    uint32_t tile_lo;            // first B tile of this launch (tile = B-half hi bits << 5 | lo bits)
    unsigned long long *err;     // += buckets with more occupied neighbour cores than entries (never expected)
};

__global__ void __launch_bounds__(256) core_present_kernel(const uint32_t *level0, uint32_t n_buckets, uint32_t *present) {
    const uint32_t k = blockIdx.x * 256u + threadIdx.x;
    const uint32_t w = __ballot_sync(kFull, k < n_buckets && level0[k] != 0u);
    if ((threadIdx.x & 31) == 0 && k < n_buckets) present[k >> 5] = w;
}

// One guide per occupied core stands for the core (whichever claims it first): its record is what
// the gather copies.  The other guides of a shared core (5 % of 100 k guides) are listed as extras.
__global__ void __launch_bounds__(256) core_claim_kernel(const uint2 *guides, uint32_t n_guides, int cl, int dl, uint32_t *claim,
                                                         uint2 *core_rec, uint32_t *extras, uint32_t *n_extras) {
    const uint32_t g = blockIdx.x * 256u + threadIdx.x;
    if (g >= n_guides) return;
    const uint2 G = __ldg(&guides[g]);
    const uint32_t cmask = (1u << cl) - 1, dmask = (1u << dl) - 1;
    const uint32_t key = (((G.x >> dl) & cmask) << cl) | ((G.y >> dl) & cmask);
    if (atomicAdd(&claim[key], 1u) == 0u) {
        const uint32_t dh = G.x & dmask, dlo = G.y & dmask;
        core_rec[key] = make_uint2(spread_bits(dlo) | (spread_bits(dh) << 1), g);
    } else {
        extras[atomicAdd(n_extras, 1u)] = g;
    }
}

// The extras are filed the round-1 way -- one warp per guide over the variant list, one returning
// atomic per entry -- but from the END of every bucket backwards (back[] starts at the bucket ends),
// behind the entries the gather writes from the front.
__global__ void __launch_bounds__(256) idx_extras_kernel(const IdxArgs a, const uint32_t *extras, const uint32_t *n_extras) {
    const uint32_t cmask = (1u << a.cl) - 1, dmask = (1u << a.dl) - 1;
    const uint32_t lane = threadIdx.x & 31, warps = gridDim.x * 8u, n = *n_extras;
    uint2 *ent = reinterpret_cast<uint2 *>(a.keys);
    for (uint32_t x = blockIdx.x * 8u + (threadIdx.x >> 5); x < n; x += warps) {
        const uint32_t g = __ldg(&extras[x]);
        const uint2 G = __ldg(&a.guides[g]);
        const uint32_t ch = (G.x >> a.dl) & cmask, cl_ = (G.y >> a.dl) & cmask;
        const uint32_t letters = spread_bits(G.y & dmask) | (spread_bits(G.x & dmask) << 1);
        for (uint32_t v = lane; v < a.n_variants; v += 32) {
            const uint32_t var = __ldg(&a.variants[v]);
            const uint32_t xh = var & 0xffff, xl = var >> 16;
            const uint32_t key = ((ch ^ xh) << a.cl) | (cl_ ^ xl);
            if (a.part_on && (key >> a.part_shift) != a.part_ix) continue;
            const uint32_t pos = atomicSub(&a.back[key], 1u) - 1u;
            ent[pos] = make_uint2(letters | (min((uint32_t)__popc(xh | xl) + a.cbias, 15u) << 20), g);
        }
    }
}

__global__ void __launch_bounds__(kGatherWarps * 32, 5) idx_gather_kernel(const GatherArgs a) {
    __shared__ uint32_t s_rows[(kGatherMaxRows + 1) * 32];
    __shared__ uint32_t s_var[kGatherMaxVar];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile = a.tile_lo + blockIdx.x / kGatherSub, sub = blockIdx.x % kGatherSub;
    const uint32_t hiB = tile >> 5, loB = tile & 31u;
    for (uint32_t x = threadIdx.x; x < (kGatherMaxRows + 1) * 32; x += kGatherWarps * 32) {
        const uint32_t row = x >> 5, hiA = x & 31u;
        uint32_t w = 0;
        if (row < a.n_rows) {
            const uint32_t rm = __ldg(&a.row_mask[row]);
            w = __ldg(&a.present[((hiB ^ (rm >> 5)) << 10) | (hiA << 5) | (loB ^ (rm & 31u))]);
        }
        // swizzled (the probes of a round hit different banks) and bit-reversed (a rotate by the
        // probe's bit index brings its bit to the top, where one funnel shift files it)
        s_rows[(row << 5) | (hiA ^ (row & 31u))] = __brev(w);
    }
    for (uint32_t x = threadIdx.x; x < (uint32_t)kGatherMaxVar; x += kGatherWarps * 32) s_var[x] = __ldg(&a.var_mask[x]);
    uint32_t pk[kGatherPerLane];
#pragma unroll
    for (int r = 0; r < kGatherPerLane; r++) pk[r] = __ldg(&a.probe[lane * kGatherPerLane + r]);
    __syncthreads();
    const uint32_t *my_var = s_var + lane * kGatherPerLane;
    constexpr uint32_t kPerCta = 1024 / kGatherSub;
    const uint32_t key_fixed = (hiB << 15) | (loB << 5);
    unsigned long long bad = 0;
    // bucket b of the CTA: A half = hi plane bits << 5 | lo plane bits
    uint32_t abits = sub * kPerCta + warp;
    uint32_t key = key_fixed | ((abits >> 5) << 10) | (abits & 31u);
    uint32_t v0 = __ldg(&a.ent_off[key]), v1 = __ldg(&a.ent_off[key + 1]);
    for (uint32_t b = warp; b < kPerCta; b += kGatherWarps) {
        const uint32_t start = v0 & ~1u, size = (v1 & ~1u) - start - (v0 & 1u);
        const uint32_t abits_now = abits, key_now = key;
        abits += kGatherWarps;
        key = key_fixed | ((abits >> 5) << 10) | (abits & 31u);
        if (b + kGatherWarps < kPerCta) {  // the next bucket's bounds travel while this one is gathered
            v0 = __ldg(&a.ent_off[key]);
            v1 = __ldg(&a.ent_off[key + 1]);
        }
        uint32_t found = 0;  // probe r ends up at bit 13 - r
#pragma unroll
        for (int r = 0; r < kGatherPerLane; r++) {
            const uint32_t t = pk[r] ^ abits_now;
            const uint32_t w = s_rows[t >> 5];
            found = __funnelshift_l(__funnelshift_l(w, w, t), found, 1);  // rotate by t & 31, then shift the top bit in
        }
        // place of the lane's first entry = occupied neighbours of the lower lanes (variant order)
        const uint32_t n = (uint32_t)__popc(found);
        uint32_t inc = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, inc, d);
            if (lane >= (uint32_t)d) inc += t;
        }
        if (lane == 31 && inc > size) bad++;
        uint2 *dst = a.ent + start + (inc - n);
        // the records of the occupied cores -> the bucket; up to four loads in flight per lane
        while (found) {
            uint2 rec[4];
            uint32_t bias[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                rec[i] = make_uint2(0u, 0u); bias[i] = 0xffffffffu;
                if (found) {
                    const uint32_t bit = 31u - (uint32_t)__clz((int)found);
                    found ^= 1u << bit;
                    const uint32_t m = my_var[(kGatherPerLane - 1) - bit];
                    rec[i] = __ldg(&a.core_rec[key_now ^ (m & 0xfffffu)]);
                    bias[i] = m & 0xf00000u;
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (bias[i] != 0xffffffffu) *dst++ = make_uint2(rec[i].x | bias[i], rec[i].y);
        }
    }
    if (bad) atomicAdd(a.err, bad);
}

// second and third stream of a handle (+ their events), created on first use
static int ensure_side_streams(DeviceState *ds) {
    if (ds->build_stream) return GUIDO_OK;
    int lo_prio = 0, hi_prio = 0;
    cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio);
    GUIDO_CUDA(cudaStreamCreateWithPriority(&ds->build_stream, cudaStreamNonBlocking, hi_prio));
    GUIDO_CUDA(cudaStreamCreateWithFlags(&ds->join_stream, cudaStreamNonBlocking));
    for (auto &e : ds->ev_slice) GUIDO_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return GUIDO_OK;
}

static void enumerate_variants(int cl, int cm, std::vector<uint32_t> *out) {
    // all (xh, xl) with popc(xh | xl) <= cm over cl positions; per chosen position 3 substitutions
    out->clear();
    std::function<void(int, uint32_t, uint32_t, int)> rec = [&](int from, uint32_t xh, uint32_t xl, int left) {
        out->push_back(xh | (xl << 16));
        if (!left) return;
        for (int p = from; p < cl; p++)
            for (int d = 1; d <= 3; d++)
                rec(p + 1, xh | ((uint32_t)(d >> 1) << p), xl | ((uint32_t)(d & 1) << p), left - 1);
    };
    rec(0, 0, 0, cm);
}

bool ot_index_applicable(const OtParams &op, int64_t n_guides) {
    if (op.raw_mode || op.core_len < 6 || op.core_len > 12 || op.core_max > 3) return false;
    if (kProto - op.core_len + op.core_max > 16) return false;  // two key fields must fit 32 bits
    std::vector<uint32_t> v;
    enumerate_variants(op.core_len, op.core_max, &v);
    return (double)v.size() * (double)n_guides < 1.5e9;  // entries are addressed with 32 bits (12 GB)
}

// builds the guide-side seed index of this search in ds->d_bucket / ds->d_index.  `entries`: one
// 8-byte {letters | bias, guide} entry per key, buckets padded to even sizes (the block join); otherwise the
// key / guide-id arrays of the popc kernels with buckets padded to 8 keys.
// `part_bits` / `part_ix`: only the buckets whose top part_bits key bits equal part_ix are built (the
// share of the key space this handle's site table covers); 0 / 0 = everything
static int build_guide_index(DeviceState *ds, const OtParams &op, const uint2 *d_guides, int64_t n_guides,
                             bool entries, int part_bits, int part_ix, cudaStream_t stream, int *n_launches,
                             OtIndex *out, const std::function<int(uint32_t, uint32_t)> *after_slice = nullptr,
                             int min_slice_bits = 0) {
    if (!ot_index_applicable(op, n_guides)) {
        set_error("seed index needs 6 <= core_length <= 12, core_mismatches <= 3 and < 1.5e9 index entries");
        return GUIDO_ERR_ARG;
    }
    const int cl = op.core_len, dl = kProto - cl;
    std::vector<uint32_t> variants;
    enumerate_variants(cl, op.core_max, &variants);
    const uint32_t n_buckets = 1u << (2 * cl);
    const uint64_t n_entries = (uint64_t)variants.size() * (uint64_t)n_guides;
    const uint32_t n_scan = n_buckets + 1, n_scan_blocks = (n_scan + 4095) / 4096;
    // d_bucket: [counts/offsets n_buckets+1][fill n_buckets][back n_buckets][block sums][variants][levels]
    // bucket sizes by the Hamming-ball convolution (ball_count_kernel) when a half core fits its
    // shared-memory tile; GUIDO_IDX_HIST=atomic forces the histogram pass (tests, comparisons)
    const char *hist_mode = getenv("GUIDO_IDX_HIST");
    const bool dp_counts = cl <= 10 && op.core_max >= 1 && op.core_max <= 3 && !(hist_mode && hist_mode[0] == 'a');
    const int64_t level_words = dp_counts ? (int64_t)(op.core_max + 1) * n_buckets : 0;
    // entries by gather (idx_gather_kernel) instead of the atomic scatter: 10-base cores, <= 2 core
    // mismatches (448 probe slots); GUIDO_IDX_BUILD=scatter forces the scatter (tests, comparisons)
    const char *build_mode = getenv("GUIDO_IDX_BUILD");
    const bool gather = entries && dp_counts && cl == 10 && variants.size() <= (size_t)kGatherMaxVar &&
                        !(build_mode && build_mode[0] == 's');
    // gather scratch behind the levels: [present n_buckets/32][probe][var_mask][row_mask 128][extras n_guides + counter][core_rec 2 n_buckets]
    const int64_t gather_words = gather ? (int64_t)n_buckets / 32 + 2 * kGatherMaxVar + 128 + n_guides + 4 + 2 * (int64_t)n_buckets : 0;
    const int64_t bucket_words = (int64_t)n_scan + 2 * (int64_t)n_buckets + n_scan_blocks + (int64_t)variants.size() + 16 + level_words +
                                 gather_words;
    int rc;
    if ((rc = ensure_bytes(&ds->d_bucket, &ds->bucket_cap, bucket_words * 4))) return rc;
    // popc format: every bucket is padded to a multiple of 8 slots; worst case allocated
    const uint32_t pad = entries ? 2u : 8u;  // entries: buckets start 16-byte aligned (bulk copies)
    const uint64_t n_slots = (n_entries + (uint64_t)(pad ? pad - 1 : 0) * n_buckets + 31) & ~(uint64_t)31;
    if (n_slots >= (1ull << 32)) { set_error("seed index too large"); return GUIDO_ERR_ARG; }
    // d_index: [keys n_slots][ids n_slots], or n_slots 8-byte entries
    if ((rc = ensure_bytes(&ds->d_index, &ds->index_cap, (int64_t)n_slots * 8))) return rc;
    uint32_t *counts = (uint32_t *)ds->d_bucket, *fill = counts + n_scan, *back = fill + n_buckets, *sums = back + n_buckets,
             *d_var = sums + n_scan_blocks;
    uint32_t *keys = (uint32_t *)ds->d_index, *ids = keys + n_slots;
    const int field = dl + op.core_max;  // <= 16, checked by ot_index_applicable
    GUIDO_CUDA(cudaMemsetAsync(counts, 0, ((size_t)n_scan + n_buckets) * 4, stream));
    // The fill scatters 8-byte entries all over a ~0.35 GB index: done in one go every store dirties
    // its own 32-byte sector in DRAM (4x write amplification).  Filling one slice of the bucket
    // range at a time -- the buckets that share their top slice_bits key bits -- keeps the slice
    // L2-resident, so DRAM sees whole lines.  Expected slots actually used (uniform keys: half a
    // pad unit per bucket) size the slices.
    const uint64_t used_slots = std::min<uint64_t>(n_slots, n_entries + (uint64_t)(pad / 2) * n_buckets);
    const char *smb = getenv("GUIDO_IDX_SLICE_MB");  // tuning knob
    const uint64_t slice_bytes = (uint64_t)(smb ? std::max(1, atoi(smb)) : 64) << 20;
    int slice_bits = std::min(5, std::max(part_bits, min_slice_bits));  // a partition is a union of slices
    // (the gather writes whole lines in bucket order and needs no slices: one launch, then one join)
    while (!gather && slice_bits < 5 && slice_bits < cl && ((used_slots * 8) >> slice_bits) > slice_bytes) slice_bits++;
    // gather tables.  Lane l owns the probe slots 14 l .. 14 l + 13 and the compacted list follows slot
    // order, so the variants of the largest core-mismatch class (405 of 436) take the slots of lanes
    // 0 .. 28 and the others follow: the block join meets the classes one after the other.  WHICH
    // slot of its class a variant takes is free, and round r probes the slots {14 l + r}: they are
    // dealt so that the 32 probes of a round fall into different shared-memory banks or onto the
    // same word (a row is stored with its word index XORed by row & 31, see idx_gather_kernel).
    std::vector<uint32_t> g_probe(kGatherMaxVar), g_var(kGatherMaxVar, 0u), g_row;
    if (gather) {
        auto cmm_of = [](uint32_t var) { return __builtin_popcount((var & 0xffffu) | (var >> 16)); };
        struct Slot { uint32_t probe, var; };
        std::vector<std::vector<Slot>> pool(2);  // [0]: largest class, [1]: the rest in class order
        std::vector<uint32_t> by_class(variants);
        std::stable_sort(by_class.begin(), by_class.end(), [&](uint32_t x, uint32_t y) { return cmm_of(x) > cmm_of(y); });
        for (uint32_t v : by_class) {
            const uint32_t xh = v & 0xffffu, xl = v >> 16;
            const uint32_t rm = ((xh >> 5) << 5) | (xl >> 5);  // B half: hi mask << 5 | lo mask
            size_t row = std::find(g_row.begin(), g_row.end(), rm) - g_row.begin();
            if (row == g_row.size()) g_row.push_back(rm);
            const uint32_t probe = ((uint32_t)row << 10) | (((xh & 31u) ^ ((uint32_t)row & 31u)) << 5) | (xl & 31u);
            const uint32_t var = (xh << 10) | xl | (std::min<uint32_t>((uint32_t)cmm_of(v) + (uint32_t)(7 - op.total_max), 15u) << 20);
            pool[cmm_of(v) == op.core_max ? 0 : 1].push_back(Slot{probe, var});
        }
        if (g_row.size() > (size_t)kGatherMaxRows) { set_error("internal: gather rows"); return GUIDO_ERR_ARG; }
        const uint32_t zero_row = (uint32_t)g_row.size();
        // lanes of the first pool: enough whole lanes for it; the second pool starts on the next lane
        const int lanes0 = (int)((pool[0].size() + kGatherPerLane - 1) / kGatherPerLane);
        if ((size_t)(32 - lanes0) * kGatherPerLane < pool[1].size()) { set_error("internal: gather slots"); return GUIDO_ERR_ARG; }
        std::vector<int> bank_word(kGatherPerLane * 32, -1);   // per round and bank: the word probed there
        std::vector<int> fill0(kGatherPerLane, 0), fill1(kGatherPerLane, 0);
        for (auto &x : g_probe) x = 0xffffffffu;
        auto place = [&](const Slot &sl, int which, bool allow_conflict) {
            const int word = (int)(sl.probe >> 5), bank = word & 31;
            const int lane_lo = which ? lanes0 : 0, lane_n = which ? 32 - lanes0 : lanes0;
            std::vector<int> &fill = which ? fill1 : fill0;
            for (int r = 0; r < kGatherPerLane; r++) {
                if (fill[r] >= lane_n) continue;
                int &bw = bank_word[r * 32 + bank];
                if (bw != -1 && bw != word && !allow_conflict) continue;
                if (bw == -1) bw = word;
                const int slot = (lane_lo + fill[r]++) * kGatherPerLane + r;
                g_probe[slot] = sl.probe; g_var[slot] = sl.var;
                return true;
            }
            return false;
        };
        for (int which = 0; which < 2; which++) {
            std::vector<Slot> left;
            for (const Slot &sl : pool[which]) if (!place(sl, which, false)) left.push_back(sl);
            for (const Slot &sl : left) place(sl, which, true);
        }
        // unused slots probe the all-zero row, at a word whose bank is free in their round
        for (int slot = 0; slot < kGatherMaxVar; slot++) {
            if (g_probe[slot] != 0xffffffffu) continue;
            const int r = slot % kGatherPerLane;
            int bank = 0;
            for (int k = 0; k < 32; k++) if (bank_word[r * 32 + k] == -1 || bank_word[r * 32 + k] == (int)(zero_row * 32 + k)) { bank = k; break; }
            bank_word[r * 32 + bank] = (int)(zero_row * 32 + bank);
            g_probe[slot] = (zero_row << 10) | ((uint32_t)bank << 5);
        }
    }
    // variants grouped by the top slice_bits of their hi-plane mask (see idx_build_kernel)
    IdxArgs ia;
    {
        const uint32_t n_groups = 1u << slice_bits;
        auto group_of = [&](uint32_t var) { return ((var & 0xffffu) >> (cl - slice_bits)) & (n_groups - 1u); };
        std::stable_sort(variants.begin(), variants.end(), [&](uint32_t x, uint32_t y) { return group_of(x) < group_of(y); });
        uint32_t at = 0;
        for (uint32_t grp = 0; grp <= n_groups; grp++) {
            while (at < variants.size() && group_of(variants[at]) < grp) at++;
            ia.group_off[grp] = (uint16_t)at;
        }
        for (uint32_t grp = n_groups + 1; grp < 33; grp++) ia.group_off[grp] = (uint16_t)variants.size();
    }
    // variants are tiny; a pageable async copy is staged by the driver before it returns
    GUIDO_CUDA(cudaMemcpyAsync(d_var, variants.data(), variants.size() * 4, cudaMemcpyHostToDevice, stream));
    ia.guides = d_guides; ia.n_guides = (uint32_t)n_guides; ia.variants = d_var; ia.n_variants = (uint32_t)variants.size();
    ia.slice_bits = (uint32_t)slice_bits; ia.slice = 0;
    ia.part_shift = (uint32_t)(2 * cl - part_bits); ia.part_ix = (uint32_t)part_ix; ia.part_on = part_bits > 0;
    ia.cl = cl; ia.dl = dl; ia.field = field; ia.counts = counts; ia.fill = fill; ia.keys = keys; ia.ids = ids;
    ia.pad_mask = pad ? pad - 1 : 0; ia.entries = entries ? 1 : 0; ia.cbias = entries ? (uint32_t)(7 - op.total_max) : 0u;
    ia.core_max = (uint32_t)op.core_max; ia.back = back;
    const int bgrid = (int)std::min<uint64_t>(((uint64_t)n_guides + 7) / 8, (uint64_t)ds->sm_count * 8);
    int n_count_launches = 1;
    if (dp_counts) {
        uint32_t *levels = d_var + ((variants.size() + 15) & ~(size_t)15);
        const int h = cl / 2;
        GUIDO_CUDA(cudaMemsetAsync(levels, 0, (size_t)n_buckets * 4, stream));
        core_hist_kernel<<<(unsigned)((n_guides + 255) / 256), 256, 0, stream>>>(d_guides, (uint32_t)n_guides, cl, dl, levels);
        ball_count_kernel<false><<<1u << (2 * (cl - h)), 256, 0, stream>>>(levels, n_buckets, cl, h, op.core_max, nullptr, false, 0, 0);
        ball_count_kernel<true><<<1u << (2 * h), 256, 0, stream>>>(levels, n_buckets, cl, h, op.core_max, counts, ia.part_on,
                                                                   ia.part_shift, ia.part_ix);
        n_count_launches = 3;
    } else {
        idx_build_kernel<false><<<bgrid, 256, 0, stream>>>(ia);
    }
    scan_blocks_kernel<<<n_scan_blocks, 1024, 0, stream>>>(counts, n_scan, sums, pad);
    scan_sums_kernel<<<1, 1024, 0, stream>>>(sums, n_scan_blocks);
    scan_fixup_kernel<<<n_scan_blocks, 1024, 0, stream>>>(counts, n_scan, sums, gather ? nullptr : fill, ia.pad_mask);
    if (entries && !gather) idx_back_cursor_kernel<<<(n_buckets + 255) / 256, 256, 0, stream>>>(counts, n_buckets, ia.pad_mask, back);
    GatherArgs ga;
    bool extras_on_side = false;
    if (gather) {
        // one guide per occupied core is claimed as the core's record; the others (cores shared by
        // several guides) are scattered from the bucket ends backwards
        uint32_t *levels = d_var + ((variants.size() + 15) & ~(size_t)15);
        uint32_t *present = levels + level_words, *d_probe = present + n_buckets / 32, *d_gvar = d_probe + kGatherMaxVar,
                 *d_row = d_gvar + kGatherMaxVar;
        uint32_t *extras = d_row + 128, *n_extras = extras + n_guides;
        uint2 *core_rec = reinterpret_cast<uint2 *>(n_extras + 2 + (((uintptr_t)(n_extras + 2) >> 2) & 1u));
        uint32_t *claim = fill;  // n_buckets words, free in gather mode
        GUIDO_CUDA(cudaMemcpyAsync(d_probe, g_probe.data(), kGatherMaxVar * 4, cudaMemcpyHostToDevice, stream));
        GUIDO_CUDA(cudaMemcpyAsync(d_gvar, g_var.data(), kGatherMaxVar * 4, cudaMemcpyHostToDevice, stream));
        GUIDO_CUDA(cudaMemcpyAsync(d_row, g_row.data(), g_row.size() * 4, cudaMemcpyHostToDevice, stream));
        GUIDO_CUDA(cudaMemsetAsync(claim, 0, (size_t)n_buckets * 4, stream));
        GUIDO_CUDA(cudaMemsetAsync(n_extras, 0, 4, stream));
        core_present_kernel<<<(n_buckets + 255) / 256, 256, 0, stream>>>(levels, n_buckets, present);
        core_claim_kernel<<<(unsigned)((n_guides + 255) / 256), 256, 0, stream>>>(d_guides, (uint32_t)n_guides, cl, dl, claim, core_rec,
                                                                                  extras, n_extras);
        idx_back_cursor_kernel<<<(n_buckets + 255) / 256, 256, 0, stream>>>(counts, n_buckets, ia.pad_mask, back);
        // the extras are few (one warp per guide, a chain of 14 atomics each): they run on a side stream
        // next to the gather, which writes the fronts of the buckets; the join waits for both
        if ((rc = ensure_side_streams(ds))) return rc;
        cudaStream_t side = stream == ds->build_stream ? ds->join_stream : ds->build_stream;
        GUIDO_CUDA(cudaEventRecord(ds->ev_slice[34], stream));
        GUIDO_CUDA(cudaStreamWaitEvent(side, ds->ev_slice[34], 0));
        idx_extras_kernel<<<bgrid, 256, 0, side>>>(ia, extras, n_extras);
        GUIDO_CUDA(cudaEventRecord(ds->ev_slice[35], side));
        extras_on_side = true;
        ga.present = present; ga.core_rec = core_rec; ga.ent_off = counts;
        ga.ent = reinterpret_cast<uint2 *>(keys); ga.probe = d_probe; ga.var_mask = d_gvar; ga.row_mask = d_row;
        ga.n_rows = (uint32_t)g_row.size(); ga.tile_lo = 0; ga.err = ds->d_counts + 3;
        if (n_launches) *n_launches += 4;
    }
    // walk buckets warp-wide when they are large; GUIDO_OT_WIDE=0/1 overrides (tuning, tests)
    const char *force = getenv("GUIDO_OT_WIDE");
    // 2 = one thread per window (default), 1 = one warp per window, 0 = flattened warp walk
    const int wide = force ? atoi(force) : 2;
    *out = OtIndex{counts, keys, ids, cl, dl, field, wide};
    // Slice by slice; the slot ranges are read from the scanned offsets on the device.  After each
    // slice the caller may start consuming it (`after_slice`: the join of that key range).
    const uint32_t per_part = 1u << (slice_bits - part_bits);
    const int key_shift = 2 * cl - slice_bits;
    for (uint32_t p = (uint32_t)part_ix * per_part; p < ((uint32_t)part_ix + 1) * per_part; p++) {
        const uint32_t key_lo = p << key_shift, key_hi = (p + 1) << key_shift;
        ia.slice = p;
        if (gather) {
            // the buckets of a slice share the top bits of their key = of the B tile index
            const uint32_t tiles = 1024u >> slice_bits;
            ga.tile_lo = p * tiles;
            idx_gather_kernel<<<tiles * kGatherSub, kGatherWarps * 32, 0, stream>>>(ga);
            if (extras_on_side) {  // (once: the extras cover every slice)
                GUIDO_CUDA(cudaStreamWaitEvent(stream, ds->ev_slice[35], 0));
                extras_on_side = false;
            }
        } else {
            idx_build_kernel<true><<<bgrid, 256, 0, stream>>>(ia);
        }
        GUIDO_CUDA(cudaGetLastError());
        if (after_slice && (rc = (*after_slice)(key_lo, key_hi))) return rc;
    }
    if (n_launches) *n_launches += 3 + (entries ? 1 : 0) + n_count_launches + (int)per_part;
    return GUIDO_OK;
}

int launch_ot_index(DeviceState *ds, const PlanesView &pv, const OtParams &op, const ScanRange &r,
                    const uint2 *d_guides, int64_t n_guides, guido_ot_hit *d_hits, uint64_t cap,
                    unsigned long long *d_count, unsigned long long *d_site_count,
                    cudaStream_t stream, int *n_launches) {
    if (r.hi <= r.lo || n_guides == 0) return GUIDO_OK;
    OtScanArgs a;
    int rc = build_guide_index(ds, op, d_guides, n_guides, false, 0, 0, stream, n_launches, &a.ix);
    if (rc) return rc;
    a.pv = pv; a.op = op; a.r = r;
    a.guides = d_guides; a.n_guides = (uint32_t)n_guides;
    a.hits = d_hits; a.cap = cap; a.count = d_count; a.flags = nullptr; a.site_count = d_site_count;
    a.tab_slots = nullptr; a.tab_counts = a.tab_fill = nullptr; a.tab_n = 0;
    rc = launch_ot_kernel<kModeIndex>(ds, a, stream);
    if (rc) return rc;
    if (n_launches) *n_launches += 1;
    return GUIDO_OK;
}

// ------------------------------------------------------------------------- PAM-site table
//
// Genome-side index: the PAM-valid windows of the resident shard in guide orientation, grouped by
// core key (counting sort: histogram pass, scan, scatter pass -- both passes are the search front
// end with a different consumer).  Every bucket is padded to whole blocks of 64 slots.  Per slot 16
// bytes {position, hi plane | strand, lo plane, core key} -- read for hits, and by the popc join --
// plus, for 10-base cores, 5 bytes of the bit-sliced block form the block join streams
// (join_blocks.cu): ~8.4 GB for NGG on 3.1 Gb.  Built once per (PAM sets, core length) and kept
// until the image is released or other parameters are asked for.  It plays the role of the
// reference's bowtie index files (genome.py:157-168): guide-independent, part of the built genome.

void release_site_table(DeviceState *ds) {
    if (ds->d_site_slots) cudaFree(ds->d_site_slots);
    if (ds->d_site_blocks) cudaFree(ds->d_site_blocks);
    ds->d_site_slots = nullptr;
    ds->d_site_blocks = nullptr;
    ds->site_cap = 0;
    ds->site_slots = 0;
    ds->site_n = -1;
}

static bool site_table_matches(const DeviceState *ds, const OtParams &op) {
    if (ds->site_n < 0 || ds->site_core_len != op.core_len || ds->site_pam_len != op.pam.len) return false;
    if (ds->site_tab_part_ix != ds->site_part_ix || ds->site_tab_n_parts != ds->site_n_parts) return false;
    for (int k = 0; k < op.pam.len; k++)
        if (ds->site_fw[k] != op.pam.fw[k] || ds->site_rv[k] != op.pam.rv[k]) return false;
    return true;
}

// bits of the core key that select the partition (n_parts is a power of two, checked when it is set)
static int part_bits_of(const DeviceState *ds) {
    int b = 0;
    while ((1 << b) < ds->site_n_parts) b++;
    return b;
}

// GUIDO_ERR_CAPACITY (without touching the last error text) when a table cannot be held: more than
// 2^31 windows or not enough free device memory; callers fall back to the streaming search
int ensure_site_table(DeviceState *ds, const PlanesView &pv, const OtParams &op, const ScanRange &r,
                      cudaStream_t stream, bool *built) {
    if (built) *built = false;
    if (site_table_matches(ds, op)) return GUIDO_OK;
    if (op.raw_mode || op.core_len < 6 || op.core_len > 12) return GUIDO_ERR_CAPACITY;
    ds->site_n = -1;
    const int cl = op.core_len, dl = kProto - cl;
    const uint32_t n_buckets = 1u << (2 * cl);
    const uint32_t n_scan = n_buckets + 1, n_scan_blocks = (n_scan + 4095) / 4096;
    int rc;
    if ((rc = ensure_bytes(&ds->d_bucket, &ds->bucket_cap, ((int64_t)n_buckets + n_scan_blocks + 16) * 4))) return rc;
    // the bucket offsets belong to the table (the joins read them), the fill cursors are scratch
    if (ds->d_site_off) { cudaFree(ds->d_site_off); ds->d_site_off = nullptr; }
    GUIDO_CUDA(cudaMalloc(&ds->d_site_off, (size_t)n_scan * 4));
    if (!ds->d_tickets) GUIDO_CUDA(cudaMalloc(&ds->d_tickets, 64 * sizeof(uint32_t)));
    uint32_t *counts = ds->d_site_off, *fill = (uint32_t *)ds->d_bucket, *sums = fill + n_buckets;
    GUIDO_CUDA(cudaMemsetAsync(counts, 0, (size_t)n_scan * 4, stream));
    GUIDO_CUDA(cudaMemsetAsync(fill, 0, (size_t)n_buckets * 4, stream));
    GUIDO_CUDA(cudaMemsetAsync(ds->d_counts + 4, 0, 2 * sizeof(unsigned long long), stream));
    GUIDO_CUDA(cudaEventRecord(ds->ev[1], stream));
    OtScanArgs a;
    a.pv = pv; a.op = op; a.r = r;
    a.ix = OtIndex{nullptr, nullptr, nullptr, cl, dl, 0, 0};
    a.guides = nullptr; a.n_guides = 0;
    a.hits = nullptr; a.cap = 0; a.count = nullptr; a.flags = nullptr;
    a.site_count = ds->d_counts + 4;  // every PAM-valid window of the shard, 64-bit: guards the 32-bit scan
    a.tab_slots = nullptr; a.tab_counts = counts; a.tab_fill = fill; a.tab_n = 0;
    a.tab_key_span = n_buckets >> part_bits_of(ds);
    a.tab_key_lo = a.tab_key_span * (uint32_t)ds->site_part_ix;
    if ((rc = launch_ot_kernel<kModeTabCount>(ds, a, stream))) return rc;
    // A 32-bit scan of the padded counts: refuse before it can wrap (every bucket gains < 64 slots)
    unsigned long long seen[2] = {0, 0};  // windows of the shard, windows in this handle's key range
    GUIDO_CUDA(cudaMemcpyAsync(seen, ds->d_counts + 4, sizeof seen, cudaMemcpyDeviceToHost, stream));
    GUIDO_CUDA(cudaStreamSynchronize(stream));
    const unsigned long long all_windows = seen[1];
    if (seen[0] >= (1ull << 31)) return GUIDO_ERR_CAPACITY;  // 32-bit slot numbers (and the scan) end here
    // bucket starts, every bucket rounded up to whole blocks of 64 slots (the pad count rides in the
    // low 6 bits); the entry behind the last bucket is the number of slots
    scan_blocks_kernel<<<n_scan_blocks, 1024, 0, stream>>>(counts, n_scan, sums, kSiteBlockSlots);
    scan_sums_kernel<<<1, 1024, 0, stream>>>(sums, n_scan_blocks);
    scan_fixup_kernel<<<n_scan_blocks, 1024, 0, stream>>>(counts, n_scan, sums, nullptr, kSiteBlockSlots - 1);
    uint32_t total = 0;
    for (int k = 0; k <= 32; k++)  // first slot of every 1/32 of the key space: the join's slice boundaries
        GUIDO_CUDA(cudaMemcpyAsync(&ds->site_row_off[k], counts + (size_t)k * (n_buckets / 32), sizeof(uint32_t),
                                   cudaMemcpyDeviceToHost, stream));
    GUIDO_CUDA(cudaMemcpyAsync(&total, counts + n_buckets, sizeof total, cudaMemcpyDeviceToHost, stream));
    GUIDO_CUDA(cudaStreamSynchronize(stream));
    for (int k = 0; k <= 32; k++) ds->site_row_off[k] &= ~(uint32_t)(kSiteBlockSlots - 1);
    total &= ~(uint32_t)(kSiteBlockSlots - 1);
    const bool want_blocks = dl == 10;  // the block join is written for ten distal bases
    if ((int64_t)total > ds->site_cap || (want_blocks && !ds->d_site_blocks)) {
        release_site_table(ds);
        size_t free_b = 0, total_b = 0;
        GUIDO_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const int64_t cap = (int64_t)total + 1024;
        const uint64_t bytes = (uint64_t)cap * kSiteSlotBytes + (want_blocks ? (uint64_t)(cap / kSiteBlockSlots) * kSiteBlockWords * 4 : 0);
        if (bytes + (4ull << 30) > free_b) return GUIDO_ERR_CAPACITY;
        // GUIDO_SITE_TABLE_MAX_MB: cap on the table size (tests of the streaming fallback, shared GPUs)
        if (const char *lim = getenv("GUIDO_SITE_TABLE_MAX_MB"))
            if (bytes > ((uint64_t)std::max(0, atoi(lim)) << 20)) return GUIDO_ERR_CAPACITY;
        GUIDO_CUDA(cudaMalloc(&ds->d_site_slots, (size_t)cap * kSiteSlotBytes));
        if (want_blocks) GUIDO_CUDA(cudaMalloc(&ds->d_site_blocks, (size_t)(cap / kSiteBlockSlots) * kSiteBlockWords * 4));
        ds->site_cap = cap;
    }
    // pad slots keep position ~0 (never a window: the image ends below 2^32 - 23)
    GUIDO_CUDA(cudaMemsetAsync(ds->d_site_slots, 0xff, (size_t)total * kSiteSlotBytes, stream));
    a.tab_slots = ds->d_site_slots;
    a.site_count = nullptr;
    if ((rc = launch_ot_kernel<kModeTabFill>(ds, a, stream))) return rc;
    if (want_blocks && (rc = launch_site_blocks(ds, ds->d_site_slots, total / kSiteBlockSlots, dl, ds->d_site_blocks, stream))) return rc;
    GUIDO_CUDA(cudaEventRecord(ds->ev[2], stream));
    GUIDO_CUDA(cudaStreamSynchronize(stream));
    cudaEventElapsedTime(&ds->site_build_ms, ds->ev[1], ds->ev[2]);
    ds->site_n = (int64_t)all_windows;
    ds->site_slots = (int64_t)total;
    ds->site_core_len = cl; ds->site_pam_len = op.pam.len;
    ds->site_tab_part_ix = ds->site_part_ix; ds->site_tab_n_parts = ds->site_n_parts;
    for (int k = 0; k < kMaxPam; k++) { ds->site_fw[k] = op.pam.fw[k]; ds->site_rv[k] = op.pam.rv[k]; }
    if (built) *built = true;
    return GUIDO_OK;
}

// The block join pays per (32 entries x 64 slots); it wins once buckets hold enough of both.
// GUIDO_OT_BLOCKS=0/1 forces the choice (tests, tuning).
static bool ot_blocks_applicable(const DeviceState *ds, const OtParams &op, int64_t n_guides) {
    if (!ds->d_site_blocks || op.core_len != 10 || op.total_max < 0 || op.total_max > 7 || op.core_max > 3) return false;
    if (const char *f = getenv("GUIDO_OT_BLOCKS")) return atoi(f) != 0;
    std::vector<uint32_t> v;
    enumerate_variants(op.core_len, op.core_max, &v);
    const double n_buckets = (double)(1u << (2 * op.core_len)) / ds->site_tab_n_parts;
    return (double)v.size() * (double)n_guides / ds->site_tab_n_parts >= 8.0 * n_buckets && (double)ds->site_n >= 64.0 * n_buckets;
}

int launch_ot_join(DeviceState *ds, const PlanesView &pv, const OtParams &op, const ScanRange &r,
                   const uint2 *d_guides, int64_t n_guides, guido_ot_hit *d_hits, uint64_t cap,
                   unsigned long long *d_count, unsigned long long *d_site_count,
                   cudaStream_t stream, int *n_launches, const OtSlices *slices) {
    const int n_out = slices && slices->n > 1 ? slices->n : 1;
    auto record_all = [&]() -> int {  // nothing (more) to search: every slice is complete
        for (int s = 0; slices && s < n_out; s++) {
            if (slices->send) {
                int rc = pack_send_device(ds, d_hits + (size_t)s * cap, d_count + s, slices->send_stride - 1,
                                          slices->send + (size_t)s * slices->send_stride, stream);
                if (rc) return rc;
            }
            if (slices->events) GUIDO_CUDA(cudaEventRecord(slices->events[s], stream));
        }
        return GUIDO_OK;
    };
    if (r.hi <= r.lo || n_guides == 0) return record_all();
    if (!site_table_matches(ds, op)) { set_error("PAM-site table was not built for these parameters"); return GUIDO_ERR_ARG; }
    if (ds->site_n == 0) return record_all();
    OtScanArgs a;
    const bool blocks = ot_blocks_applicable(ds, op, n_guides);
    a.pv = pv; a.op = op; a.r = r;
    a.guides = d_guides; a.n_guides = (uint32_t)n_guides;
    a.hits = d_hits; a.cap = cap; a.count = d_count; a.flags = nullptr; a.site_count = d_site_count;
    a.tab_slots = ds->d_site_slots;
    a.tab_counts = a.tab_fill = nullptr;
    a.first_block = 0; a.n_tiles = 0;
    int per_sm = 0;
    GUIDO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ot_join_kernel, kThreads, 0));
    const int max_grid = ds->sm_count * std::max(per_sm, 1) * 4;  // 4 waves' worth of CTAs evens out the tail
    if (d_site_count) {  // statistics: windows of the table (the kernels add the entries they compare)
        const unsigned long long n_windows = (unsigned long long)ds->site_n;
        GUIDO_CUDA(cudaMemcpyAsync(d_site_count, &n_windows, sizeof n_windows, cudaMemcpyHostToDevice, stream));
    }
    BjArgs bj;
    bj.blocks = ds->d_site_blocks; bj.blk_off = ds->d_site_off; bj.slots = ds->d_site_slots;
    bj.ent = nullptr; bj.ent_off = nullptr; bj.key_lo = bj.key_hi = 0; bj.group = bj.n_groups = 0;
    bj.ticket = ds->d_tickets; bj.hits = d_hits; bj.cap = cap; bj.count = d_count; bj.site_count = d_site_count;
    const int cl = op.core_len;
    const int part_bits = part_bits_of(ds);
    const uint32_t part_keys = (1u << (2 * cl)) >> part_bits;
    const uint32_t part_lo = part_keys * (uint32_t)ds->site_part_ix;
    if (blocks) GUIDO_CUDA(cudaMemsetAsync(ds->d_tickets, 0, 64 * sizeof(uint32_t), stream));

    // one join launch over the keys [key_lo, key_hi) on stream s
    int n_joins = 0;
    auto join_keys = [&](uint32_t key_lo, uint32_t key_hi, cudaStream_t s) -> int {
        const uint32_t unit = (1u << (2 * cl)) / 32;
        if (blocks) {
            bj.ent = reinterpret_cast<const uint2 *>(a.ix.keys); bj.ent_off = a.ix.offsets;
            bj.key_lo = key_lo; bj.key_hi = key_hi; bj.ticket = ds->d_tickets + (n_joins % 64);
            int rc = launch_join_blocks(ds, bj, s);
            if (rc) return rc;
        } else {
            a.tab_lo = ds->site_row_off[key_lo / unit];
            a.tab_n = ds->site_row_off[key_hi / unit];
            if (a.tab_n <= a.tab_lo) return GUIDO_OK;
            const uint32_t n_chunks = (a.tab_n - a.tab_lo + kJoinPerThread * kThreads - 1) / (kJoinPerThread * kThreads);
            const int grid = (int)std::min<uint32_t>((uint32_t)max_grid, n_chunks);
            ot_join_kernel<<<grid, kThreads, 0, s>>>(a);
            GUIDO_CUDA(cudaGetLastError());
        }
        n_joins++;
        return GUIDO_OK;
    };

    // The index is built slice by slice (a slice = the buckets that share their top key bits); the
    // table is in key order, so as soon as a slice is ready the join of ITS buckets can start on
    // the caller's stream while the next slice is being filled on a second, high-priority stream:
    // the scatter-bound build (atomics, little issue pressure) hides behind the ALU-bound join.
    // GUIDO_OT_OVERLAP=0 (measurement knob): everything on the caller's stream, ONE join launch over
    // the whole table after the index is complete -- guido_kernel_times then reports the join
    // kernel timed alone (bench.py's roofline pass); the results are the same either way.
    if (slices && blocks && (n_out & (n_out - 1)) == 0 && part_bits + __builtin_ctz((unsigned)n_out) <= 5) {
        // sliced output: index and join slice after slice on the caller's stream; a slice's hits, count and
        // event are complete before the next slice starts
        int slice_bits = part_bits + __builtin_ctz((unsigned)n_out);
        const uint32_t keys_per_out = part_keys / (uint32_t)n_out;
        const std::function<int(uint32_t, uint32_t)> join_slice = [&](uint32_t key_lo, uint32_t key_hi) -> int {
            const uint32_t s_ix = (key_lo - part_lo) / keys_per_out;
            bj.hits = d_hits + (size_t)s_ix * cap;
            bj.count = d_count + s_ix;
            if (n_joins == 0) kev_begin(ds, stream);  // guido_kernel_times: first join launch .. last join done
            int rc = join_keys(key_lo, key_hi, stream);
            if (rc) return rc;
            if ((key_hi - part_lo) % keys_per_out == 0) {  // the slice is complete
                if (slices->send && (rc = pack_send_device(ds, d_hits + (size_t)s_ix * cap, d_count + s_ix, slices->send_stride - 1,
                                                           slices->send + (size_t)s_ix * slices->send_stride, stream))) return rc;
                if (slices->events) GUIDO_CUDA(cudaEventRecord(slices->events[s_ix], stream));
            }
            return GUIDO_OK;
        };
        int rc = build_guide_index(ds, op, d_guides, n_guides, blocks, part_bits, ds->site_part_ix, stream, n_launches, &a.ix,
                                   &join_slice, slice_bits);
        if (n_joins) kev_end(ds, stream);
        if (rc) return rc;
        if (n_launches) *n_launches += n_joins;
        return GUIDO_OK;
    }
    const char *ov = getenv("GUIDO_OT_OVERLAP");
    if (slices || (ov && atoi(ov) == 0)) {  // (slices that cannot be honoured: everything lands in slice 0)
        int rc = build_guide_index(ds, op, d_guides, n_guides, blocks, part_bits, ds->site_part_ix, stream, n_launches, &a.ix);
        if (rc) return rc;
        kev_begin(ds, stream);
        rc = join_keys(part_lo, part_lo + part_keys, stream);
        kev_end(ds, stream);
        if (rc) return rc;
        if (n_launches) *n_launches += n_joins;
        return record_all();
    }
    { int rc = ensure_side_streams(ds); if (rc) return rc; }
    // joins alternate between the caller's stream and a sibling, so the tail of one slice's join
    // (its last, partly filled wave of CTAs) runs next to the head of the following one
    cudaStream_t sb = ds->build_stream, sj = ds->join_stream;
    GUIDO_CUDA(cudaEventRecord(ds->ev_slice[33], stream));        // guides are on the device, previous search is done
    GUIDO_CUDA(cudaStreamWaitEvent(sb, ds->ev_slice[33], 0));
    GUIDO_CUDA(cudaStreamWaitEvent(sj, ds->ev_slice[33], 0));
    const std::function<int(uint32_t, uint32_t)> join_slice = [&](uint32_t key_lo, uint32_t key_hi) -> int {
        const uint32_t unit = (1u << (2 * cl)) / 32;
        cudaEvent_t ready = ds->ev_slice[(key_lo / unit) % 31];
        GUIDO_CUDA(cudaEventRecord(ready, sb));
        cudaStream_t s = (n_joins & 1) ? sj : stream;
        GUIDO_CUDA(cudaStreamWaitEvent(s, ready, 0));
        if (n_joins == 0) kev_begin(ds, stream);  // guido_kernel_times: first join launch .. last join done
        return join_keys(key_lo, key_hi, s);
    };
    int rc = build_guide_index(ds, op, d_guides, n_guides, blocks, part_bits, ds->site_part_ix, sb, n_launches,
                               &a.ix, &join_slice);
    // the caller's stream continues when both join streams (and with them the build stream) are done
    GUIDO_CUDA(cudaEventRecord(ds->ev_slice[32], sj));
    GUIDO_CUDA(cudaStreamWaitEvent(stream, ds->ev_slice[32], 0));
    GUIDO_CUDA(cudaEventRecord(ds->ev_slice[31], sb));  // (slices without table rows are built but never joined)
    GUIDO_CUDA(cudaStreamWaitEvent(stream, ds->ev_slice[31], 0));
    if (rc) return rc;
    if (n_joins) kev_end(ds, stream);
    if (n_launches) *n_launches += n_joins;
    return GUIDO_OK;
}

}  // namespace guido
