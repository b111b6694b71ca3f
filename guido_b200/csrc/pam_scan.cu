// pam_scan.cu -- whole-genome / region PAM scan, both strands, ordered compaction.
//
// Sits under Locus._find_guides_in_interval (reference guido/locus.py:178-183): the two
// overlapping `re.finditer` passes become two streaming passes over the bit-plane image.
//
// One launch (pam_onepass_kernel, ranges below 2^31 bases): per tile (1024 blocks = 65536 bases) hit
// masks -> popcounts -> tile counts published in a descriptor -> hits staged in shared memory ->
// output offset from a decoupled look-back over the preceding tiles' descriptors -> coalesced
// stores.  The loop is software pipelined so that a look-back walks instead of polling (see the
// kernel).  Larger ranges use two launches without any inter-CTA waiting:
//   pam_count_kernel  per tile: hit masks -> popcounts -> tile counts
//   pam_fill_kernel   per tile: hit masks again (the planes come from L2 when the shard fits,
//                     else from HBM); the tile's output offset is the running sum of the tile
//                     counts (each CTA adds the counts of the tiles it strides over); CTA prefix
//                     sum; hits staged in shared memory; coalesced 4-byte stores
// (History of the single-pass form -- a first version polled ~1200 descriptors per tile and lost
// to count -> fill; what made it work is in profiles/r1_pam_scan.md, section F.)
//
// A thread owns 4 consecutive blocks (256 bases), handled as eight 32-bit words per plane:
//   * two 256-bit loads (LDG.E.256) fetch its 64 bytes of {lo,hi} words; the next block (the
//     PAM-length halo) comes from lane+1 by shuffle, lane 31 loads it itself;
//   * set membership per PAM position is 4 LOP3 per word, branch-free: every base set is a
//     boolean function of (hi, lo), evaluated in algebraic normal form
//         f = c0 ^ (c1 & lo) ^ (c2 & hi) ^ (c3 & hi & lo)
//     with all-ones / all-zeros coefficients prepared on the host; PAM offsets are funnel shifts;
//   * the run veto (non-ACGT runs, contig pads) happens only in the rare tiles that touch a run,
//     as bit masks: invalid-base masks per word, dilated by the window length with shift-OR
//     doubling -- no per-hit loop, no per-hit global load.
//
// HBM traffic per scan (algorithmic): 0.25 B/base read + 4 B per hit written.
#include <algorithm>

#include "scan_common.cuh"

namespace guido {

constexpr int kBpt = 4;                        // blocks per thread
constexpr int kWpt = 2 * kBpt;                 // 32-bit words per thread and plane
constexpr int kScanTileBlocks = kThreads * kBpt;
constexpr int kScanTileBases = kScanTileBlocks * 64;
constexpr int kStage = 5120;  // staged hits per strand and tile (NGG: 4096 expected, sigma 62)

static_assert(kScanTileBases == (1 << kTileRunShift), "tile_first_run granularity must match the scan tile");

// per strand and PAM position: ANF coefficients of the base set, or skip (set = all four bases)
struct PamCoef {
    uint32_t c[2][kMaxPam][4];
    uint8_t skip[2][kMaxPam];
    uint8_t same[2][kMaxPam];  // same set as the previous non-skipped position: reuse its masks
    int len;
};

struct PamScanArgs {
    PlanesView pv;
    PamCoef pc;
    ScanRange r;
    int mode;
    uint32_t first_tile;  // absolute tile index (tile t covers blocks [1024 t, 1024 t + 1024))
    uint32_t n_tiles;
    uint32_t *out_fw, *out_rv;
    unsigned long long cap_fw, cap_rv;
    unsigned long long *counts;   // [0] n_fw, [1] n_rv
    unsigned long long *tiles;    // n_tiles: per-tile counts, fw | rv << 32
    int vec_ok;                   // both output arrays are 16-byte aligned: copy out as uint4
    unsigned long long *err;      // one-pass kernel: set when a look-back gave up (never expected)
};

// planes of the thread's 4 blocks + the following block, as 32-bit words (10 per plane)
struct ThreadWords {
    uint32_t lo[kWpt + 2], hi[kWpt + 2];
};

// raw registers of one thread's tile slice: requested early (software pipelining), expanded late
struct ThreadRaw {
    unsigned long long q[2 * kBpt];  // lo0, hi0, lo1, hi1, ...
    ulonglong2 halo;                 // lane 31 only: the block after the warp's last
};

// Shared-memory staging of one tile's planes: kBpt rows of kThreads blocks (thread t owns column
// t, so its 16-byte reads and cp.async writes are conflict free) + one halo block per warp.  A
// thread reads back only what it requested itself, so waiting for its own cp.async group is all
// the synchronisation there is.  (Keeping the next tile in registers instead looked cheaper and
// was not: the compiler parked those registers in local memory, and the store waited for the load.)
constexpr int kStageBlocks = kBpt * kThreads + kThreads / 32;

__device__ __forceinline__ void cp_async16(ulonglong2 *smem_dst, const ulonglong2 *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src));
}

// request the planes of the thread's kBpt blocks starting at block0 (+ the warp's halo block)
__device__ __forceinline__ void stage_issue(const PlanesView &pv, int64_t block0, int lane, ulonglong2 *buf) {
    const int64_t i = block0 - pv.blk_lo;
    if (i >= 0 && i + kBpt <= pv.blk_n) {
#pragma unroll
        for (int k = 0; k < kBpt; k++) cp_async16(buf + k * kThreads + threadIdx.x, pv.planes + i + k);
    } else {  // tile sticks out of the resident range: zeros outside
#pragma unroll
        for (int k = 0; k < kBpt; k++) buf[k * kThreads + threadIdx.x] = load_block(pv, block0 + k);
    }
    if (lane == 31) {
        ulonglong2 *h = buf + kBpt * kThreads + (threadIdx.x >> 5);
        if (i + kBpt >= 0 && i + kBpt < pv.blk_n) cp_async16(h, pv.planes + i + kBpt);
        else *h = make_ulonglong2(0ull, 0ull);
    }
    asm volatile("cp.async.commit_group;");
}

// the thread's own requests have landed: move them to registers
__device__ __forceinline__ void stage_read(const ulonglong2 *buf, int lane, ThreadRaw &raw) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int k = 0; k < kBpt; k++) {
        const ulonglong2 b = buf[k * kThreads + threadIdx.x];
        raw.q[2 * k] = b.x; raw.q[2 * k + 1] = b.y;
    }
    raw.halo = make_ulonglong2(0ull, 0ull);
    if (lane == 31) raw.halo = buf[kBpt * kThreads + (threadIdx.x >> 5)];
}

__device__ __forceinline__ void expand_thread_words(const ThreadRaw &raw, int lane, ThreadWords &t) {
    const unsigned long long *q = raw.q;
    ulonglong2 halo = raw.halo;
    const ulonglong2 first = make_ulonglong2(q[0], q[1]);
    const ulonglong2 nb = shfl_down_u2(first, 1);
    if (lane != 31) halo = nb;
#pragma unroll
    for (int k = 0; k < kBpt; k++) {
        t.lo[2 * k] = (uint32_t)q[2 * k]; t.lo[2 * k + 1] = (uint32_t)(q[2 * k] >> 32);
        t.hi[2 * k] = (uint32_t)q[2 * k + 1]; t.hi[2 * k + 1] = (uint32_t)(q[2 * k + 1] >> 32);
    }
    t.lo[kWpt] = (uint32_t)halo.x; t.lo[kWpt + 1] = (uint32_t)(halo.x >> 32);
    t.hi[kWpt] = (uint32_t)halo.y; t.hi[kWpt + 1] = (uint32_t)(halo.y >> 32);
}

// hit words of one strand: bit i of h[w] <=> the PAM matches at base (thread base + 32 w + i)
__device__ __forceinline__ void strand_hit_words(const ThreadWords &t, const PamCoef &pc, int strand, uint32_t h[kWpt]) {
#pragma unroll
    for (int w = 0; w < kWpt; w++) h[w] = 0xffffffffu;
    uint32_t m[kWpt + 1];
    for (int i = 0; i < pc.len; i++) {
        if (pc.skip[strand][i]) continue;  // N: every base; non-ACGT bases are removed by the run veto
        if (!pc.same[strand][i]) {         // e.g. the second G of NGG reuses the first G's masks
            const uint32_t c0 = pc.c[strand][i][0], c1 = pc.c[strand][i][1], c2 = pc.c[strand][i][2], c3 = pc.c[strand][i][3];
#pragma unroll
            for (int w = 0; w < kWpt + 1; w++)
                m[w] = c0 ^ (c1 & t.lo[w]) ^ (c2 & t.hi[w]) ^ (c3 & t.hi[w] & t.lo[w]);
        }
#pragma unroll
        for (int w = 0; w < kWpt; w++) h[w] &= __funnelshift_r(m[w], m[w + 1], i);  // i < 32
    }
}

// words [a, a+L) OR-ed together for every start bit: z bit p = OR_{d<L} v bit (p+d), L <= 32
__device__ __forceinline__ uint32_t dilate(uint32_t w0, uint32_t w1, int L) {
    unsigned long long v = ((unsigned long long)w1 << 32) | w0;
    int k = 1;
    while (2 * k <= L) { v |= v >> k; k *= 2; }
    if (L > k) v |= v >> (L - k);
    return (uint32_t)v;
}

// Veto words for the thread's 8 hit words (rare path: tile touches >= 1 run).  A forward hit at p
// owns window [p-20, p+P), a reverse hit [p, p+P+20) (guides.py:53-67); what is checked depends
// on the scan mode (see GUIDO_SCAN_* in the header).
__device__ __forceinline__ void veto_words(const PlanesView &pv, int r_lo, int r_hi, int mode, int P, uint32_t base,
                                           uint32_t hf[kWpt], uint32_t hr[kWpt]) {
    // most threads of a tile that touches a run are nowhere near it: skip them (and whole warps)
    bool near = false;
    for (int r = r_lo; r < r_hi; r++)
        near |= __ldg(&pv.run_start[r]) < base + 32 * (kWpt + 2) && __ldg(&pv.run_end[r]) + 32 > base;
    if (!near) return;
    // invalid-base words -1 .. 9 around the thread's 256 bases (index w + 1)
    uint32_t any[kWpt + 3], k0[kWpt + 3];
#pragma unroll
    for (int w = 0; w < kWpt + 3; w++) any[w] = k0[w] = 0;
    for (int r = r_lo; r < r_hi; r++) {
        const int64_t s = (int64_t)__ldg(&pv.run_start[r]) - ((int64_t)base - 32);
        const int64_t e = (int64_t)__ldg(&pv.run_end[r]) - ((int64_t)base - 32);
        const bool kind0 = __ldg(&pv.run_kind[r]) == 0;
#pragma unroll
        for (int w = 0; w < kWpt + 3; w++) {
            const int64_t a = s - 32 * w, b = e - 32 * w;
            if (b <= 0 || a >= 32) continue;
            const uint32_t hi_m = b >= 32 ? 0xffffffffu : ((1u << (int)b) - 1);
            const uint32_t lo_m = a <= 0 ? 0u : ((1u << (int)a) - 1);
            const uint32_t m = hi_m & ~lo_m;
            any[w] |= m;
            if (kind0) k0[w] |= m;
        }
    }
    const int W = P + kProto;
#pragma unroll
    for (int w = 0; w < kWpt; w++) {
        // word w of the thread sits at index w + 1
        const uint32_t pam_bad = dilate(any[w + 1], any[w + 2], P);
        uint32_t fbad = pam_bad, rbad = pam_bad;
        if (mode == GUIDO_SCAN_GUIDE || mode == GUIDO_SCAN_STRICT) {
            const bool g = (mode == GUIDO_SCAN_GUIDE);  // '+' guides drop literal N / pads only
            const uint32_t s0 = g ? k0[w] : any[w], s1 = g ? k0[w + 1] : any[w + 1], s2 = g ? k0[w + 2] : any[w + 2];
            // window starts 20 bases before the PAM: z(p - 20)
            const uint32_t z0 = dilate(s0, s1, W), z1 = dilate(s1, s2, W);
            fbad |= __funnelshift_l(z0, z1, kProto);
            rbad = dilate(any[w + 1], any[w + 2], W);
        } else if (mode == kScanRawWindow) {
            fbad = rbad = dilate(any[w + 1], any[w + 2], W);
        }
        hf[w] &= ~fbad;
        hr[w] &= ~rbad;
    }
}

// restrict hit words to PAM starts p with r.lo <= p < r.hi and p + P <= r.limit (boundary tiles)
__device__ __forceinline__ uint32_t range_word(int64_t base, const ScanRange &r, int P) {
    int64_t hi = (int64_t)r.hi;
    const int64_t lim = (int64_t)r.limit - P + 1;
    if (lim < hi) hi = lim;
    int64_t a = (int64_t)r.lo - base, b = hi - base;
    if (a < 0) a = 0;
    if (b > 32) b = 32;
    if (b <= a) return 0;
    const uint32_t hm = b >= 32 ? 0xffffffffu : ((1u << (int)b) - 1);
    const uint32_t lm = a <= 0 ? 0u : ((1u << (int)a) - 1);
    return hm & ~lm;
}

// both strands' hit words of one thread for tile `tile` (all zero for tiles inside one run)
__device__ __forceinline__ int64_t tile_block0(const PamScanArgs &a, uint32_t tile) {
    return (int64_t)(a.first_tile + tile) * kScanTileBlocks + threadIdx.x * kBpt;
}

// `raw` holds the thread's planes of this tile, requested one tile earlier (stage_issue / stage_read)
__device__ __forceinline__ void tile_hit_words(const PamScanArgs &a, uint32_t tile, const ThreadRaw &raw,
                                               uint32_t hf[kWpt], uint32_t hr[kWpt]) {
    const int lane = threadIdx.x & 31;
    const int P = a.pc.len;
    const uint32_t t0 = (a.first_tile + tile) * (uint32_t)kScanTileBases;
    int r_lo, r_hi;
    bool skip;
    tile_runs(a.pv, t0, t0 + kScanTileBases, r_lo, r_hi, skip);
    const int64_t block0 = tile_block0(a, tile);
    ThreadWords t;
    expand_thread_words(raw, lane, t);
#pragma unroll
    for (int w = 0; w < kWpt; w++) hf[w] = hr[w] = 0;
    if (skip) return;
    strand_hit_words(t, a.pc, 0, hf);
    strand_hit_words(t, a.pc, 1, hr);
    const uint32_t base = (uint32_t)(block0 * 64);
    // only the first / last tile of a range can be cut by it
    const int64_t lim = min((long long)a.r.hi, (long long)a.r.limit - P + 1);
    if ((int64_t)t0 < (int64_t)a.r.lo || (int64_t)t0 + kScanTileBases > lim) {
#pragma unroll
        for (int w = 0; w < kWpt; w++) {
            const uint32_t rm = range_word((int64_t)base + 32 * w, a.r, P);
            hf[w] &= rm; hr[w] &= rm;
        }
    }
    if (r_hi > r_lo) veto_words(a.pv, r_lo, r_hi, a.mode, P, base, hf, hr);
}

__device__ __forceinline__ unsigned long long popc_words(const uint32_t hf[kWpt], const uint32_t hr[kWpt]) {
    uint32_t cf = 0, cr = 0;
#pragma unroll
    for (int w = 0; w < kWpt; w++) { cf += __popc(hf[w]); cr += __popc(hr[w]); }
    return (unsigned long long)cf | ((unsigned long long)cr << 32);
}

// sum of a packed counter over the CTA (every thread gets it); s_warp: kThreads/32 words
__device__ __forceinline__ unsigned long long cta_sum(unsigned long long v, unsigned long long *s_warp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) t += s_warp[w];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(kThreads, 4) pam_count_kernel(const PamScanArgs a) {
    // one buffer is enough: a thread moves its slots to registers (stage_read) and only then asks
    // for the next tile's planes in the same slots
    __shared__ __align__(16) ulonglong2 s_planes[kStageBlocks];
    const int lane = threadIdx.x & 31;
    if (blockIdx.x < a.n_tiles) stage_issue(a.pv, tile_block0(a, blockIdx.x), lane, s_planes);
    for (uint32_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        ThreadRaw raw;
        stage_read(s_planes, lane, raw);
        // request the next tile's planes before working on this tile's: one tile is always in flight
        if (tile + gridDim.x < a.n_tiles) stage_issue(a.pv, tile_block0(a, tile + gridDim.x), lane, s_planes);
        uint32_t hf[kWpt], hr[kWpt];
        tile_hit_words(a, tile, raw, hf, hr);
        // no CTA barrier anywhere in this kernel: warps add their counts to the (zeroed) tile slot
        unsigned long long v = popc_words(hf, hr);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        if (lane == 0 && v) atomicAdd(&a.tiles[tile], v);
    }
}

// set bits of a word -> positions base+bit appended to dst[off...]
// A 32-base word holds Poisson(2) NGG hits: the first four are extracted by predicated
// straight-line code (no loop, no divergence penalty), the rare rest by a loop.
__device__ __forceinline__ uint32_t emit_word(uint32_t m, uint32_t base, uint32_t *dst, uint32_t off) {
    uint32_t *p = dst + off;
    const uint32_t next = off + __popc(m);
    const uint32_t b1 = base - 1;  // __ffs is 1-based
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (m) p[k] = b1 + __ffs(m);
        m &= m - 1;  // 0 stays 0
    }
    uint32_t n = 4;
    while (m) { p[n++] = b1 + __ffs(m); m &= m - 1; }
    return next;
}

// The same into the shared-memory stage.  The word's slots are filled from the TOP bit down (slot
// rank = popc - 1 - k): `bfind` (FLO) finds the highest set bit in one XU instruction where
// __ffs needs two (BREV + FLO), and the XU pipe -- POPC, BREV, FLO -- was the busiest pipe of this
// kernel (ncu: 51 %).  The four straight-line stores are predicated by hand: left to itself the
// compiler wraps each of them in a divergence region (BSSY / BRA / BSYNC), which triples the
// instruction count of the emission.
__device__ __forceinline__ uint32_t emit_word_staged(uint32_t m, uint32_t base, uint32_t stage_addr, uint32_t off) {
    const uint32_t next = off + __popc(m);
    uint32_t addr = stage_addr + 4u * next;  // shared-window byte address just past the word's last slot
#pragma unroll
    for (int k = 0; k < 4; k++) {
        // top = index of the highest set bit (0xffffffff when m == 0: nothing stored, the shift
        // clamps to 0 and m stays 0)
        asm volatile("{ .reg .pred q; .reg .u32 top, pos, bit;\n\t"
                     "bfind.u32 top, %0;\n\t"
                     "setp.ne.u32 q, %0, 0;\n\t"
                     "add.u32 pos, %2, top;\n\t"
                     "@q st.shared.u32 [%1], pos;\n\t"
                     "shl.b32 bit, 1, top;\n\t"
                     "xor.b32 %0, %0, bit; }"
                     : "+r"(m) : "r"(addr - 4u * (k + 1)), "r"(base) : "memory");
    }
    addr -= 16;
    while (m) {
        const uint32_t top = 31u - (uint32_t)__clz((int)m);
        addr -= 4;
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(base + top) : "memory");
        m ^= 1u << top;
    }
    return next;
}

// Emission with a deferred tail (one-pass kernel).  A 32-base word holds Poisson(2) NGG hits, but a
// warp pays for the MAXIMUM over its lanes -- 6.1 slots per word, four of them straight-line and
// the rest in a loop that nearly every warp-word enters.  Here a word gets three straight-line
// slots; a lane whose word holds more parks the remainder {bits, where its slots end, word index}
// in a private queue and drains it once, after all sixteen words of the tile: the warp then pays
// for the maximum of the lanes' TOTAL leftovers (~10 rounds) instead of sixteen per-word maxima.
// The queue is the thread's own column of the plane buffer, free once the planes are in registers
// (8 entries of 8 bytes; when it is full the word is finished in place).
constexpr int kEmitQueue = 2 * kBpt;  // entries per thread: its kBpt 16-byte plane slots

__device__ __forceinline__ uint32_t emit_word_deferred(uint32_t m, uint32_t base, uint32_t stage_addr, uint32_t off,
                                                        uint32_t w, uint32_t q_addr, uint32_t &q_n) {
    const uint32_t next = off + __popc(m);
    uint32_t addr = stage_addr + 4u * next;  // shared-window byte address just past the word's last slot
#pragma unroll
    for (int k = 0; k < 3; k++) {
        asm volatile("{ .reg .pred q; .reg .u32 top, pos, bit;\n\t"
                     "bfind.u32 top, %0;\n\t"
                     "setp.ne.u32 q, %0, 0;\n\t"
                     "add.u32 pos, %2, top;\n\t"
                     "@q st.shared.u32 [%1], pos;\n\t"
                     "shl.b32 bit, 1, top;\n\t"
                     "xor.b32 %0, %0, bit; }"
                     : "+r"(m) : "r"(addr - 4u * (k + 1)), "r"(base) : "memory");
    }
    if (m) {
        addr -= 12;
        if (q_n < (uint32_t)kEmitQueue) {  // entry e lives in plane slot e / 2 of this thread's column, half e % 2
            const uint32_t e_addr = q_addr + (q_n >> 1) * (uint32_t)(kThreads * 16) + (q_n & 1u) * 8u;
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(e_addr), "r"(m), "r"(addr | (w << 24)) : "memory");
            q_n++;
        } else {
            while (m) {
                const uint32_t top = 31u - (uint32_t)__clz((int)m);
                addr -= 4;
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(base + top) : "memory");
                m ^= 1u << top;
            }
        }
    }
    return next;
}

// thread_base = position of the thread's first base (the words' bases are thread_base + 32 w)
__device__ __forceinline__ void emit_drain(uint32_t thread_base, uint32_t q_addr, uint32_t q_n) {
    for (uint32_t e = 0; e < q_n; e++) {
        uint32_t m, meta;
        const uint32_t e_addr = q_addr + (e >> 1) * (uint32_t)(kThreads * 16) + (e & 1u) * 8u;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(m), "=r"(meta) : "r"(e_addr) : "memory");
        uint32_t addr = meta & 0xffffffu;
        const uint32_t base = thread_base + 32u * (meta >> 24);
        while (m) {
            const uint32_t top = 31u - (uint32_t)__clz((int)m);
            addr -= 4;
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(base + top) : "memory");
            m ^= 1u << top;
        }
    }
}

// staged hits [a0, a0+n) of s -> g[a0 .. a0+n); s and g agree modulo 16 bytes, so whole uint4s
// move in the middle and at most 3 scalars at either end
__device__ __forceinline__ void copy_out(const uint32_t *s, uint32_t a0, uint32_t n, uint32_t *g, int vec_ok) {
    const uint32_t end = a0 + n;
    const uint32_t q0 = (a0 + 3) >> 2, q1 = end >> 2;
    if (vec_ok && q1 > q0) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(s);
        uint4 *g4 = reinterpret_cast<uint4 *>(g);
        for (uint32_t q = q0 + threadIdx.x; q < q1; q += kThreads) g4[q] = s4[q];
        const uint32_t h = a0 + threadIdx.x, t = 4 * q1 + threadIdx.x;
        if (h < 4 * q0) g[h] = s[h];
        if (t < end) g[t] = s[t];
    } else {
        for (uint32_t i = a0 + threadIdx.x; i < end; i += kThreads) g[i] = s[i];
    }
}

constexpr int kFillPlaneBytes = kStageBlocks * (int)sizeof(ulonglong2);
constexpr int kFillSmemBytes = kFillPlaneBytes + 2 * (kStage + 8) * (int)sizeof(uint32_t);

// (4 CTAs/SM -- 64 registers, a smaller stage -- was measured: spills, no gain)
__global__ void __launch_bounds__(kThreads, 3) pam_fill_kernel(const PamScanArgs a) {
    __shared__ unsigned long long s_warp[kThreads / 32], s_part[kThreads / 32];
    // dynamic shared memory (above the static limit): [plane buffer][2 hit stages]
    extern __shared__ __align__(16) unsigned char s_dyn[];
    ulonglong2 *s_planes = reinterpret_cast<ulonglong2 *>(s_dyn);
    uint32_t (*s_stage)[kStage + 8] = reinterpret_cast<uint32_t (*)[kStage + 8]>(s_dyn + kFillPlaneBytes);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t sa_f = (uint32_t)__cvta_generic_to_shared(s_stage[0]);
    const uint32_t sa_r = (uint32_t)__cvta_generic_to_shared(s_stage[1]);
    // running output offsets of the current tile = counts of all tiles before it
    unsigned long long ef = 0, er = 0;
    uint32_t summed = 0;  // tiles [0, summed) are in (ef, er)
    if (blockIdx.x < a.n_tiles) stage_issue(a.pv, tile_block0(a, blockIdx.x), lane, s_planes);
    for (uint32_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        ThreadRaw raw;
        stage_read(s_planes, lane, raw);
        // next tile's planes fly while this tile is scanned, staged and stored
        if (tile + gridDim.x < a.n_tiles) stage_issue(a.pv, tile_block0(a, tile + gridDim.x), lane, s_planes);
        // advance (ef, er) over the tiles [summed, tile): at most gridDim.x counts (<= 2^16 each, so
        // the packed 32-bit halves cannot overflow), L2 resident; reduced alongside the CTA scan
        unsigned long long part = 0;
        for (uint32_t u = summed + threadIdx.x; u < tile; u += kThreads) part += __ldg(&a.tiles[u]);
        summed = tile;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(kFull, part, o);
        uint32_t hf[kWpt], hr[kWpt];
        tile_hit_words(a, tile, raw, hf, hr);
        const unsigned long long mine = popc_words(hf, hr);
        // CTA exclusive scan of the packed (fw, rv) counts
        unsigned long long inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long t = __shfl_up_sync(kFull, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) { s_warp[warp] = inc; s_part[warp] = part; }
        __syncthreads();
        unsigned long long warp_prefix = 0, total = 0, adv = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; w++) {
            const unsigned long long x = s_warp[w];
            if (w < warp) warp_prefix += x;
            total += x;
            adv += s_part[w];
        }
        ef += adv & 0xffffffffull;
        er += adv >> 32;
        const unsigned long long excl = warp_prefix + inc - mine;
        const uint32_t agg_f = (uint32_t)total, agg_r = (uint32_t)(total >> 32);
        const int64_t block0 = (int64_t)(a.first_tile + tile) * kScanTileBlocks + threadIdx.x * kBpt;
        const uint32_t base = (uint32_t)(block0 * 64) - a.r.out_base;
        // staged at index (ef & 3) + local rank, so shared and global addresses share their 16-byte
        // phase and the copy-out below can move uint4s
        const uint32_t pf = (uint32_t)ef & 3u, pr = (uint32_t)er & 3u;
        uint32_t of = (uint32_t)excl + pf, orv = (uint32_t)(excl >> 32) + pr;
        // capacity: a tile is written completely or not at all (the counts stay exact either way)
        const bool fits_f = ef + agg_f <= a.cap_fw, fits_r = er + agg_r <= a.cap_rv;
        const bool stage_f = agg_f <= kStage, stage_r = agg_r <= kStage;
        if (stage_f) {
#pragma unroll
            for (int w = 0; w < kWpt; w++) of = emit_word_staged(hf[w], base + 32 * w, sa_f, of);
        } else if (fits_f) {  // unusually dense tile: straight from registers to HBM
#pragma unroll
            for (int w = 0; w < kWpt; w++) of = emit_word(hf[w], base + 32 * w, a.out_fw + ef - pf, of);
        }
        if (stage_r) {
#pragma unroll
            for (int w = 0; w < kWpt; w++) orv = emit_word_staged(hr[w], base + 32 * w, sa_r, orv);
        } else if (fits_r) {
#pragma unroll
            for (int w = 0; w < kWpt; w++) orv = emit_word(hr[w], base + 32 * w, a.out_rv + er - pr, orv);
        }
        __syncthreads();
        if (stage_f && fits_f) copy_out(s_stage[0], pf, agg_f, a.out_fw + ef - pf, a.vec_ok);
        if (stage_r && fits_r) copy_out(s_stage[1], pr, agg_r, a.out_rv + er - pr, a.vec_ok);
        if (tile == a.n_tiles - 1 && threadIdx.x == 0) {
            a.counts[0] = ef + agg_f;
            a.counts[1] = er + agg_r;
        }
        __syncthreads();  // s_stage / s_warp are rewritten by the next iteration
    }
}

// ---- one pass: tile counts and output offsets in the same kernel (decoupled look-back) ----------
//
// The count kernel costs a quarter of the scan and reads the planes a second time.  Here a tile is
// scanned once: its CTA publishes the tile's hit counts as soon as they are known, warp 0 walks
// back over the descriptors of the preceding tiles until it meets an inclusive prefix -- WHILE the
// other warps already stage the tile's hits in shared memory, which only needs tile-local ranks --
// and the offsets are first needed for the copy-out.  Tiles are handed out by an atomic ticket at
// the moment a CTA is ready to start one, so every predecessor of a running tile is running or
// done (no deadlock, and nobody waits for a tile that is merely reserved).
//
// descriptor = status << 62 | fw hits << 31 | rv hits; status 0 = nothing yet, 1 = the tile's own
// counts, 2 = inclusive prefix.  31 bits per strand: launch_pam_scan takes this path only for
// ranges below 2^31 bases.  One aligned 64-bit word, relaxed loads / stores: the word IS the data.
constexpr unsigned long long kDescAgg = 1ull << 62, kDescPrefix = 2ull << 62, kDescValue = (1ull << 62) - 1;
constexpr uint32_t kLookbackSpins = 1u << 24;  // polls before a look-back gives up (seconds: a bug, never a load effect)

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// whole warp: packed exclusive prefix (fw << 31 | rv) of `tile` from the descriptors before it.
// A round reads 256 descriptors (8 per lane, independent loads, one L2 round trip): tiles are
// started at ~40 per microsecond across the GPU, so a walk of 32 descriptors per round trip never
// catches up with the tiles that have published their counts but not yet their prefix.
constexpr int kLookPerLane = 8;
__device__ __forceinline__ unsigned long long look_back(const PamScanArgs &a, uint32_t tile, int lane) {
    unsigned long long excl = 0;
    int64_t base = (int64_t)tile - 1;
    uint32_t spins = 0;
    if (ld_relaxed_u64(a.err)) return 0;  // an earlier look-back gave up: the launch is lost, do not wait again
    while (base >= 0) {
        unsigned long long d[kLookPerLane];
#pragma unroll
        for (int k = 0; k < kLookPerLane; k++) {  // distance lane + 32 k: k = 0, lane 0 is the nearest predecessor
            const int64_t idx = base - lane - 32 * k;
            d[k] = idx >= 0 ? ld_relaxed_u64(&a.tiles[idx]) : kDescPrefix;  // before tile 0: prefix 0
        }
        unsigned long long v = 0;
        bool done = false, retry = false;
#pragma unroll
        for (int k = 0; k < kLookPerLane; k++) {
            if (!done && !retry) {  // warp-uniform
                const uint32_t st = (uint32_t)(d[k] >> 62);
                const uint32_t pmask = __ballot_sync(kFull, st == 2), imask = __ballot_sync(kFull, st == 0);
                // lanes up to and including the nearest inclusive prefix (all 32 when there is none)
                const uint32_t upto = pmask ? ((2u << (__ffs((int)pmask) - 1)) - 1u) : 0xffffffffu;
                if (imask & upto) retry = true;  // a predecessor has not published yet: poll again
                else {
                    if ((upto >> lane) & 1u) v += d[k] & kDescValue;
                    done = pmask != 0;
                }
            }
        }
        if (retry) {
            if (++spins > kLookbackSpins) { if (lane == 0) atomicExch(a.err, 1ull); break; }
            continue;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        excl += v;
        if (done) break;
        base -= 32 * kLookPerLane;
    }
    return excl;
}

// staged hits s[0, n) -> g[0, n) (g = output + tile offset, any 4-byte phase): whole uint4 stores
// where g is 16-byte aligned; the shared side is read as two aligned uint4s and shifted by the
// (CTA-uniform) phase
__device__ __forceinline__ void copy_out_shifted(const uint32_t *s, uint32_t n, uint32_t *g) {
    const uint32_t d0 = (uint32_t)((0u - (uint32_t)(reinterpret_cast<uintptr_t>(g) >> 2)) & 3u);  // elements up to alignment
    if (n < d0 + 4u) {
        for (uint32_t i = threadIdx.x; i < n; i += kThreads) g[i] = s[i];
        return;
    }
    const uint32_t nq = (n - d0) >> 2;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(s);
    uint4 *g4 = reinterpret_cast<uint4 *>(g + d0);
    if (d0 == 0) {
        for (uint32_t q = threadIdx.x; q < nq; q += kThreads) g4[q] = s4[q];
    } else if (d0 == 1) {
        for (uint32_t q = threadIdx.x; q < nq; q += kThreads) { const uint4 x = s4[q], y = s4[q + 1]; g4[q] = make_uint4(x.y, x.z, x.w, y.x); }
    } else if (d0 == 2) {
        for (uint32_t q = threadIdx.x; q < nq; q += kThreads) { const uint4 x = s4[q], y = s4[q + 1]; g4[q] = make_uint4(x.z, x.w, y.x, y.y); }
    } else {
        for (uint32_t q = threadIdx.x; q < nq; q += kThreads) { const uint4 x = s4[q], y = s4[q + 1]; g4[q] = make_uint4(x.w, y.x, y.y, y.z); }
    }
    if (threadIdx.x < d0) g[threadIdx.x] = s[threadIdx.x];
    const uint32_t t = d0 + 4u * nq + threadIdx.x;
    if (t < n) g[t] = s[t];
}

// Software pipeline over the CTA's tiles: the COUNT of tile n+1 (load, hit words, scan, counts
// published) runs before the look-back and copy-out of tile n, whose hits wait in the stage.  A
// look-back therefore starts a whole phase after its own tile's counts went out, when the tiles
// just before it -- started at almost the same moment, but with a long-tailed time to their counts
// (N-run veto, DRAM latency) -- have long published theirs: it walks (1-2 rounds) and never polls.
// Measured with %globaltimer stamps: ticket -> counts 3.5 us (p99 8.4), look-back right after the
// tile's own staging 3.3 us, of which > 2 us was waiting for the slowest of ~300 predecessors.
__global__ void __launch_bounds__(kThreads, 3) pam_onepass_kernel(const PamScanArgs a) {
    __shared__ unsigned long long s_warp[kThreads / 32];
    __shared__ unsigned long long s_excl;
    __shared__ uint32_t s_ticket;
    extern __shared__ __align__(16) unsigned char s_dyn[];
    ulonglong2 *s_planes = reinterpret_cast<ulonglong2 *>(s_dyn);
    uint32_t (*s_stage)[kStage + 8] = reinterpret_cast<uint32_t (*)[kStage + 8]>(s_dyn + kFillPlaneBytes);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t sa_f = (uint32_t)__cvta_generic_to_shared(s_stage[0]);
    const uint32_t sa_r = (uint32_t)__cvta_generic_to_shared(s_stage[1]);
    unsigned long long *ticket = a.tiles + a.n_tiles;
    if (threadIdx.x == 0) s_ticket = (uint32_t)atomicAdd(ticket, 1ull);
    __syncthreads();
    uint32_t tile = s_ticket;
    // the pending tile: staged (or, if dense, already written), waiting for its offsets and copy-out
    bool pend = false, p_stage_f = false, p_stage_r = false, p_resolved = false;
    uint32_t p_tile = 0, p_agg_f = 0, p_agg_r = 0;
    unsigned long long p_ex = 0;

    // warp 0: exclusive prefix of tile t (own counts `agg`) -> s_excl; inclusive prefix published
    auto resolve_prefix = [&](uint32_t t, unsigned long long agg) {
        const unsigned long long ex = t ? look_back(a, t, lane) : 0ull;
        if (lane == 0) {
            if (t) st_relaxed_u64(&a.tiles[t], kDescPrefix | (ex + agg));
            s_excl = ex;
        }
    };

    for (;;) {
        const bool have = tile < a.n_tiles;  // CTA-uniform
        if (!have && !pend) break;
        uint32_t hf[kWpt], hr[kWpt];
        unsigned long long mine = 0, inc = 0;
        if (have) {
            stage_issue(a.pv, tile_block0(a, tile), lane, s_planes);
            {   // tickets run ahead of this tile by about one grid: pull that tile's planes into L2 (the
                // next tile of THIS CTA is not known yet, so its planes cannot be requested into shared memory)
                const int64_t i = tile_block0(a, tile + gridDim.x) - a.pv.blk_lo;
                if (tile + gridDim.x < a.n_tiles && i >= 0 && i + kBpt <= a.pv.blk_n)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a.pv.planes + i));
            }
            ThreadRaw raw;
            stage_read(s_planes, lane, raw);
            tile_hit_words(a, tile, raw, hf, hr);
            mine = popc_words(hf, hr);  // fw | rv << 32
            inc = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned long long t = __shfl_up_sync(kFull, inc, d);
                if (lane >= d) inc += t;
            }
            if (lane == 31) s_warp[warp] = inc;
        }
        __syncthreads();  // (1) warp totals are there; every warp has finished staging the pending tile
        unsigned long long warp_prefix = 0, total = 0;
        if (have) {
#pragma unroll
            for (int w = 0; w < kThreads / 32; w++) {
                const unsigned long long x = s_warp[w];
                if (w < warp) warp_prefix += x;
                total += x;
            }
        }
        const uint32_t agg_f = (uint32_t)total, agg_r = (uint32_t)(total >> 32);
        const unsigned long long agg = ((unsigned long long)agg_f << 31) | agg_r;
        if (have && threadIdx.x == 0) st_relaxed_u64(&a.tiles[tile], (tile ? kDescAgg : kDescPrefix) | agg);

        if (pend) {  // offsets and copy-out of the tile staged one iteration ago
            if (!p_resolved) {
                if (warp == 0) resolve_prefix(p_tile, ((unsigned long long)p_agg_f << 31) | p_agg_r);
                __syncthreads();  // (2)
                p_ex = s_excl;
            }
            const unsigned long long ef = p_ex >> 31, er = p_ex & 0x7fffffffull;
            // capacity: a tile is written completely or not at all (the counts stay exact either way)
            if (p_stage_f && ef + p_agg_f <= a.cap_fw) copy_out_shifted(s_stage[0], p_agg_f, a.out_fw + ef);
            if (p_stage_r && er + p_agg_r <= a.cap_rv) copy_out_shifted(s_stage[1], p_agg_r, a.out_rv + er);
            if (p_tile == a.n_tiles - 1 && threadIdx.x == 0) {
                a.counts[0] = ef + p_agg_f;
                a.counts[1] = er + p_agg_r;
            }
            __syncthreads();  // (3) the stage is free again (and s_excl has been read)
            pend = false;
        }
        if (!have) continue;  // (ends the loop: nothing pending any more)

        // the next ticket travels while this tile's hits are staged
        if (threadIdx.x == 0) s_ticket = (uint32_t)atomicAdd(ticket, 1ull);
        const unsigned long long excl = warp_prefix + inc - mine;
        const int64_t block0 = (int64_t)(a.first_tile + tile) * kScanTileBlocks + threadIdx.x * kBpt;
        const uint32_t base = (uint32_t)(block0 * 64) - a.r.out_base;
        uint32_t of = (uint32_t)excl, orv = (uint32_t)(excl >> 32);
        const bool stage_f = agg_f <= kStage, stage_r = agg_r <= kStage;
        const bool direct = !(stage_f && stage_r);  // unusually dense tile: written straight from registers,
        if (direct) {                                // so its offsets are needed now
            if (warp == 0) resolve_prefix(tile, agg);
            __syncthreads();
            p_ex = s_excl;
        }
        // (the plane buffer is free: stage_read moved this tile's planes to registers, the next tile's
        // are requested at the top of the next iteration, by each thread into its own column)
        const uint32_t q_addr = (uint32_t)__cvta_generic_to_shared(s_planes + threadIdx.x);
        uint32_t q_n = 0;
        if (stage_f) {
#pragma unroll
            for (int w = 0; w < kWpt; w++) of = emit_word_deferred(hf[w], base + 32 * w, sa_f, of, (uint32_t)w, q_addr, q_n);
            emit_drain(base, q_addr, q_n);
            q_n = 0;
        }
        if (stage_r) {
#pragma unroll
            for (int w = 0; w < kWpt; w++) orv = emit_word_deferred(hr[w], base + 32 * w, sa_r, orv, (uint32_t)w, q_addr, q_n);
            emit_drain(base, q_addr, q_n);
        }
        if (direct) {
            const unsigned long long ef = p_ex >> 31, er = p_ex & 0x7fffffffull;
            if (!stage_f && ef + agg_f <= a.cap_fw) {
#pragma unroll
                for (int w = 0; w < kWpt; w++) of = emit_word(hf[w], base + 32 * w, a.out_fw + ef, of);
            }
            if (!stage_r && er + agg_r <= a.cap_rv) {
#pragma unroll
                for (int w = 0; w < kWpt; w++) orv = emit_word(hr[w], base + 32 * w, a.out_rv + er, orv);
            }
        }
        pend = true; p_tile = tile; p_agg_f = agg_f; p_agg_r = agg_r;
        p_stage_f = stage_f; p_stage_r = stage_r; p_resolved = direct;
        __syncthreads();  // (4) the ticket is visible (s_warp / the plane buffer are free as well)
        tile = s_ticket;
    }
}

static void make_coef(const PamSpec &pam, PamCoef *pc) {
    pc->len = pam.len;
    for (int strand = 0; strand < 2; strand++) {
        int prev = -1;
        for (int i = 0; i < pam.len; i++) {
            const int s = strand ? pam.rv[i] : pam.fw[i];
            pc->same[strand][i] = (s != 15 && s == prev);
            if (s != 15) prev = s;
            // truth table over (hi, lo): A=(0,0) bit0, C=(0,1) bit1, G=(1,0) bit2, T=(1,1) bit3
            const int f00 = s & 1, f01 = (s >> 1) & 1, f10 = (s >> 2) & 1, f11 = (s >> 3) & 1;
            const int c0 = f00, c1 = f00 ^ f01, c2 = f00 ^ f10, c3 = f00 ^ f01 ^ f10 ^ f11;
            pc->c[strand][i][0] = c0 ? 0xffffffffu : 0u;
            pc->c[strand][i][1] = c1 ? 0xffffffffu : 0u;
            pc->c[strand][i][2] = c2 ? 0xffffffffu : 0u;
            pc->c[strand][i][3] = c3 ? 0xffffffffu : 0u;
            pc->skip[strand][i] = (s == 15);
        }
    }
}

int launch_pam_scan(DeviceState *ds, const PlanesView &pv, const PamSpec &pam, int mode,
                    const ScanRange &r, uint32_t *d_fw, uint64_t cap_fw, uint32_t *d_rv,
                    uint64_t cap_rv, unsigned long long *d_counts, cudaStream_t stream,
                    int *n_launches) {
    if (r.hi <= r.lo) {
        GUIDO_CUDA(cudaMemsetAsync(d_counts, 0, 2 * sizeof(unsigned long long), stream));
        return GUIDO_OK;
    }
    PamScanArgs a;
    a.pv = pv; a.r = r; a.mode = mode;
    make_coef(pam, &a.pc);
    a.first_tile = (uint32_t)((r.lo / 64) / kScanTileBlocks);
    const uint32_t last_tile = (uint32_t)((((int64_t)r.hi - 1) / 64) / kScanTileBlocks);
    a.n_tiles = last_tile - a.first_tile + 1;
    a.out_fw = d_fw; a.out_rv = d_rv; a.cap_fw = cap_fw; a.cap_rv = cap_rv;
    a.counts = d_counts;
    a.vec_ok = ((reinterpret_cast<uintptr_t>(d_fw) | reinterpret_cast<uintptr_t>(d_rv)) & 15) == 0;
    int64_t need = (int64_t)a.n_tiles + 1;  // + the ticket counter of the one-pass kernel
    if (need > ds->desc_cap) {
        if (ds->d_desc) cudaFree(ds->d_desc);
        ds->d_desc = nullptr; ds->desc_cap = 0;
        GUIDO_CUDA(cudaMalloc(&ds->d_desc, (size_t)need * sizeof(unsigned long long)));
        ds->desc_cap = need;
    }
    a.tiles = ds->d_desc;
    a.err = ds->d_counts + 7;
    if (ds->scan_grid[0] == 0) {
        int per_sm = 0;
        GUIDO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pam_count_kernel, kThreads, 0));
        ds->scan_grid[0] = ds->sm_count * std::max(per_sm, 1);
        GUIDO_CUDA(cudaFuncSetAttribute(pam_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFillSmemBytes));
        GUIDO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pam_fill_kernel, kThreads, kFillSmemBytes));
        ds->scan_grid[1] = ds->sm_count * std::max(per_sm, 1);
        GUIDO_CUDA(cudaFuncSetAttribute(pam_onepass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFillSmemBytes));
        GUIDO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pam_onepass_kernel, kThreads, kFillSmemBytes));
        ds->scan_grid[2] = ds->sm_count * std::max(per_sm, 1);
    }
    GUIDO_CUDA(cudaMemsetAsync(a.tiles, 0, (size_t)(a.n_tiles + 1) * sizeof(unsigned long long), stream));
    GUIDO_CUDA(cudaMemsetAsync(a.err, 0, sizeof(unsigned long long), stream));  // per launch; the host calls read it back
    // One pass (look-back) when the descriptors' 31-bit counters cannot overflow and there are
    // enough tiles to matter; count -> fill otherwise.  GUIDO_PAM_PASSES=1/2 forces either (tests).
    const char *force = getenv("GUIDO_PAM_PASSES");
    const bool small = (uint64_t)a.n_tiles * kScanTileBases < (1ull << 31);
    const bool one_pass = force ? (atoi(force) == 1 && small) : small;
    kev_begin(ds, stream);
    if (one_pass) {
        // every CTA of the grid must be resident: a tile only ever waits for tiles that are running
        pam_onepass_kernel<<<(int)std::min<int64_t>(a.n_tiles, ds->scan_grid[2]), kThreads, kFillSmemBytes, stream>>>(a);
    } else {
        pam_count_kernel<<<(int)std::min<int64_t>(a.n_tiles, ds->scan_grid[0]), kThreads, 0, stream>>>(a);
        pam_fill_kernel<<<(int)std::min<int64_t>(a.n_tiles, ds->scan_grid[1]), kThreads, kFillSmemBytes, stream>>>(a);
    }
    kev_end(ds, stream);
    GUIDO_CUDA(cudaGetLastError());
    if (n_launches) *n_launches += one_pass ? 1 : 2;
    return GUIDO_OK;
}

}  // namespace guido
