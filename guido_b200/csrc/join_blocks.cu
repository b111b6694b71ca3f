// join_blocks.cu -- the off-target join for large guide sets: guide-side seed index (entries in
// bucket order) x genome-side PAM-site table in BIT-SLICED BLOCKS OF 64 WINDOWS.
//
// Sits under run_bowtie (reference guido/off_targets.py:89-227), like offtarget.cu.
//
// Both sides are grouped by core k-mer (offtarget.cu): a bucket of the site table holds the
// PAM-valid windows with that core, the same bucket of the seed index the guides whose core is
// within the seed budget of it.  What is left per bucket is W windows x K entries distal compares
// (human genome, 100 k guides: W ~ 366, K ~ 42).  The table is guide independent, so IT is the side
// that is bit-sliced, once, when it is built: a block of 64 windows stores, per distal base j and
// letter c, the 64-bit word "window i differs from c at base j" (bit i) -- 40 words, 320 bytes,
// 5 bytes per window.  Buckets are padded to whole blocks (366 -> 384: 5 % pad slots instead of the
// 35 % of the round-1 layout that sliced the ~42 keys of a bucket into 64 slots).
//
// One LANE owns one index entry (a guide filed under this core): its ten letters select ten words
// of a block -- ten 8-byte shared-memory loads at per-lane addresses fixed for the whole bucket --
// and a carry-save tree of LOP3s says for which of the 64 windows (distal mismatches + bias)
// reaches 8, bias = core mismatches of the entry + 7 - total_max: no XOR/OR stage, no POPC, no
// per-window set-up.  18 LOP3 per 32 windows in general; 14 when every lane of the pass has bias 5
// ("at most 2 of 10", the dominant class: the index files those entries first in every bucket).
// Entries beyond a multiple of 32 do not get a whole pass of their own: 16 or fewer left-over
// entries share the lanes 2 or 4 ways and each copy takes every 2nd / 4th block.
//
// Data movement: a WARP streams its buckets by itself.  Lane 0 fetches the blocks and the entries
// of the next bucket with two bulk-TMA copies (cp.async.bulk -> UBLKCP, completion on the warp's own
// mbarrier) into the other half of a double buffer while the warp compares the current one; no
// CTA barrier anywhere in the loop.  Buckets are handed out in groups by an atomic ticket.
// Non-empty result words are queued per warp (a vote gives every lane that has one its slot: one
// predicated 16-byte store, no shared-memory atomic, no divergent region) and turned into hit
// records by all 32 lanes, 32 at a time: one atomic on the global cursor per round, the gathers
// (guide id, window slot) go out behind it.
#include "scan_common.cuh"
#include "join_blocks.cuh"

namespace guido {

namespace {

#ifndef BJ_WARPS
#define BJ_WARPS 8
#endif
#ifndef BJ_MIN_CTAS
#define BJ_MIN_CTAS 3
#endif
#ifndef BJ_ROLL
#define BJ_ROLL 0
#endif
#if BJ_ROLL
#define BJ_UNROLL_PRAGMA _Pragma("unroll 1")
#else
#define BJ_UNROLL_PRAGMA _Pragma("unroll")
#endif
constexpr int kBjWarps = BJ_WARPS;               // warps per CTA (they do not interact)
constexpr int kBjCapBlocks = 7;                  // 64-window blocks per stage buffer (larger buckets go in pieces)
constexpr int kBjCapEnt = 64;                    // entries per stage buffer (the rest of a bucket is read from global memory)
constexpr int kBjBlockWords = kSiteBlockWords;   // 80 32-bit words = 40 x 64 bits
constexpr int kBjQueue = 64;                     // result words parked per warp
constexpr int kBjGroupMax = 16;                  // buckets per ticket, at most
constexpr uint32_t kBjBiasAtMost2 = 5;           // bias of "at most 2 distal mismatches"

struct __align__(16) BjStage {
    uint2 blk[kBjCapBlocks * kBjBlockWords / 2]; // 2240 bytes, filled by a bulk copy
    uint2 ent[kBjCapEnt];                        // 512 bytes, filled by a bulk copy
};

struct __align__(16) BjWarpSmem {
    BjStage st[2];
    uint4 meta[kBjGroupMax];                     // per bucket of the ticket: {first block, blocks, first entry, entries}
    uint4 q[kBjQueue];                           // parked results: {hit bits of windows 0..31, 32..63, guide, block}
    unsigned long long bar[2];                   // mbarriers: "stage s is full"
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

#define BJ_LOP3(name, lut)                                                                         \
    __device__ __forceinline__ uint32_t name(uint32_t a, uint32_t b, uint32_t c) {                 \
        uint32_t r; asm("lop3.b32 %0, %1, %2, %3, " #lut ";" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; \
    }
BJ_LOP3(lop_xor3, 0x96)      // a ^ b ^ c
BJ_LOP3(lop_maj, 0xe8)       // at least two of three
BJ_LOP3(lop_nor3, 0x01)      // ~(a | b | c)
BJ_LOP3(lop_or3, 0xfe)       // a | b | c
BJ_LOP3(lop_and3, 0x80)      // a & b & c
BJ_LOP3(lop_a_and_borc, 0xe0)  // a & (b | c)
BJ_LOP3(lop_ab_or_c, 0xea)     // (a & b) | c
#undef BJ_LOP3

// what one lane keeps for the whole bucket: the ten word addresses of ITS letters in block 0 of the
// stage buffer and the four bias words (bit planes of the entry's bias, 0 or ~0)
struct LaneKey {
    const uint2 *p[10];
    uint32_t q1, q2, q4, q8;
};

// bit i set <=> window i is a hit for this lane's entry: carry-save count of the ten mismatch words
// plus the bias; only "reaches 8" matters, so sums that feed nothing are never formed (18 LOP3)
__device__ __forceinline__ uint32_t tree_any(const uint32_t (&m)[10], const LaneKey &k) {
    // weight 1: eleven inputs -> five carries
    const uint32_t s0 = lop_xor3(m[0], m[1], m[2]), c0 = lop_maj(m[0], m[1], m[2]);
    const uint32_t s1 = lop_xor3(m[3], m[4], m[5]), c1 = lop_maj(m[3], m[4], m[5]);
    const uint32_t s2 = lop_xor3(m[6], m[7], m[8]), c2 = lop_maj(m[6], m[7], m[8]);
    const uint32_t s3 = lop_xor3(s0, s1, s2), c3 = lop_maj(s0, s1, s2);
    const uint32_t c4 = lop_maj(s3, m[9], k.q1);
    // weight 2: five carries + bias bit 1 -> three carries
    const uint32_t t0 = lop_xor3(c0, c1, c2), d0 = lop_maj(c0, c1, c2);
    const uint32_t t1 = lop_xor3(c3, c4, k.q2), d1 = lop_maj(c3, c4, k.q2);
    const uint32_t d2 = t0 & t1;
    // weight 4: three carries + bias bit 2 -> two carries
    const uint32_t u0 = lop_xor3(d0, d1, d2), e0 = lop_maj(d0, d1, d2);
    const uint32_t e1 = u0 & k.q4;
    // weight 8: any carry or bias bit 3 = more than 7
    return lop_nor3(e0, e1, k.q8);
}

// the same for bias 5, i.e. "at most two of the ten words are set": with A = s0 + s1 + s2 + m9 and
// C = c0 + c1 + c2 the count is A + 2 C, and it reaches 3 when C >= 2, or C = 1 and A >= 1, or
// A >= 3 (14 LOP3)
__device__ __forceinline__ uint32_t tree_at_most2(const uint32_t (&m)[10]) {
    const uint32_t s0 = lop_xor3(m[0], m[1], m[2]), c0 = lop_maj(m[0], m[1], m[2]);
    const uint32_t s1 = lop_xor3(m[3], m[4], m[5]), c1 = lop_maj(m[3], m[4], m[5]);
    const uint32_t s2 = lop_xor3(m[6], m[7], m[8]), c2 = lop_maj(m[6], m[7], m[8]);
    const uint32_t c_ge2 = lop_maj(c0, c1, c2), c_ge1 = lop_or3(c0, c1, c2);
    const uint32_t s_any = lop_or3(s0, s1, s2), s_ge2 = lop_maj(s0, s1, s2), s_all = lop_and3(s0, s1, s2);
    const uint32_t x = lop_a_and_borc(c_ge1, s_any, m[9]);   // C >= 1 and A >= 1
    const uint32_t y = lop_ab_or_c(s_ge2, m[9], s_all);      // A >= 3
    return lop_nor3(c_ge2, x, y);
}

// hit bits of one parked result -> records: the guide id comes from the index entry, position and
// planes of each hit window from its slot of the table
__device__ __forceinline__ void bj_write_hits(const BjArgs &a, uint32_t r, uint32_t guide, size_t slot0, unsigned long long &dst) {
    while (r) {
        const uint32_t bit = (uint32_t)__ffs((int)r) - 1u;
        r &= r - 1u;
        const uint4 s = __ldg(&a.slots[slot0 + bit]);
        if (dst < a.cap) reinterpret_cast<uint4 *>(a.hits)[dst] = make_uint4(guide, s.x, s.y, s.z);
        dst++;
    }
}

// parked results -> hit records, lane i takes entry i; only whole rounds of 32 unless `all`.  The
// cursor is moved first and the gathers go out behind it, so a round costs one memory round trip.
// Whatever stays is moved to the front of the queue.
__device__ __noinline__ uint32_t bj_flush(const BjArgs &a, BjWarpSmem &ws, uint32_t qn, bool all) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t done = 0;
    while (done + 32u <= qn || (all && done < qn)) {
        const uint32_t i = done + lane;
        const uint4 v = i < qn ? ws.q[i] : make_uint4(0, 0, 0, 0);
        const uint32_t n = (uint32_t)(__popc(v.x) + __popc(v.y));
        const uint32_t total = __reduce_add_sync(kFull, n);
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.count, (unsigned long long)total);
        uint32_t guide = 0;
        uint4 s0 = make_uint4(0, 0, 0, 0);
        uint32_t r0 = v.x, r1 = v.y;
        if (n) {  // the common case is one hit in the result
            guide = v.z;
            const uint32_t first = r0 ? (uint32_t)__ffs((int)r0) - 1u : 32u + (uint32_t)__ffs((int)r1) - 1u;
            s0 = __ldg(&a.slots[(size_t)v.w * 64u + first]);
            if (r0) r0 &= r0 - 1u; else r1 &= r1 - 1u;
        }
        uint32_t inc = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, inc, d);
            if (lane >= (uint32_t)d) inc += t;
        }
        base = __shfl_sync(kFull, base, 0);
        unsigned long long dst = base + (inc - n);
        if (n) {
            if (dst < a.cap) reinterpret_cast<uint4 *>(a.hits)[dst] = make_uint4(guide, s0.x, s0.y, s0.z);
            dst++;
            bj_write_hits(a, r0, guide, (size_t)v.w * 64u, dst);
            bj_write_hits(a, r1, guide, (size_t)v.w * 64u + 32u, dst);
        }
        done += 32u;
    }
    const uint32_t left = done < qn ? qn - done : 0u;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (lane < left) v = ws.q[done + lane];
    __syncwarp();
    if (lane < left) ws.q[lane] = v;
    __syncwarp();
    return left;
}

// one pass of <= 32 entries (R of them valid) over the nblk blocks of the stage buffer; with
// S = 2 / 4 the lanes hold S copies of <= 16 / 8 entries and copy s takes blocks s, s + S, ...
// `e_first`: global index of the pass's first entry, `e_rel` its index inside the bucket.
template <int S>
__device__ __forceinline__ void bj_pass(const BjArgs &a, BjWarpSmem &ws, const BjStage &st, uint32_t nblk,
                                        uint32_t gblk0, uint32_t e_first, uint32_t e_rel, uint32_t R, uint32_t &qn) {
    constexpr uint32_t kPer = 32 / S;
    constexpr int kIters = (kBjCapBlocks + S - 1) / S;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r_ix = lane & (kPer - 1), s_ix = lane / kPer;
    uint32_t lt;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
    const bool valid = r_ix < R;
    const uint32_t e = e_first + (valid ? r_ix : 0u);
    // the lane's entry: letters | bias, and the guide it stands for (kept for the hit records: the
    // flush does not have to fetch it again from global memory)
    uint2 ent = make_uint2(15u << 20, 0u);  // bias 15 never hits
    if (valid) ent = e_rel + r_ix < (uint32_t)kBjCapEnt ? st.ent[e_rel + r_ix] : __ldg(&a.ent[e]);
    const uint32_t w = ent.x;
    LaneKey k;
#pragma unroll
    for (int j = 0; j < 10; j++) k.p[j] = st.blk + s_ix * (kBjBlockWords / 2) + ((w >> (2 * j)) & 3u);
    k.q1 = 0u - ((w >> 20) & 1u); k.q2 = 0u - ((w >> 21) & 1u); k.q4 = 0u - ((w >> 22) & 1u); k.q8 = 0u - ((w >> 23) & 1u);
    // the short tree needs bias 5 in every lane (idle lanes carry 15 and must stay silent: they
    // take the general tree's path too)
    const bool at_most2 = S == 1 && __all_sync(kFull, ((w >> 20) & 15u) == kBjBiasAtMost2);
    auto step = [&](const int it, const bool short_tree) {
        uint32_t mx[10], my[10];
#pragma unroll
        for (int j = 0; j < 10; j++) {
            const uint2 m = k.p[j][it * S * (kBjBlockWords / 2) + 4 * j];
            mx[j] = m.x; my[j] = m.y;
        }
        uint32_t rx = short_tree ? tree_at_most2(mx) : tree_any(mx, k);
        uint32_t ry = short_tree ? tree_at_most2(my) : tree_any(my, k);
        if (S > 1 && it * S + s_ix >= nblk) rx = ry = 0;  // this copy's block is beyond the piece
        // non-empty results are parked in the warp's queue: slot = queue length (a warp-uniform
        // register) + rank among the lanes that have one -- a vote and a predicated store, no
        // shared-memory atomic and no divergent region; whole rounds of 32 leave at once
        const bool has = (rx | ry) != 0u;
        const uint32_t hm = __ballot_sync(kFull, has);
        if (hm) {
            if (has) ws.q[qn + (uint32_t)__popc(hm & lt)] = make_uint4(rx, ry, ent.y, gblk0 + it * S + s_ix);
            qn += (uint32_t)__popc(hm);
            if (qn >= 32u) { __syncwarp(); qn = bj_flush(a, ws, qn, false); }
        }
    };
    if (at_most2) {
        BJ_UNROLL_PRAGMA
        for (int it = 0; it < kIters; it++) {
            if ((uint32_t)(it * S) >= nblk) break;
            step(it, true);
        }
    } else {
        BJ_UNROLL_PRAGMA
        for (int it = 0; it < kIters; it++) {
            if ((uint32_t)(it * S) >= nblk) break;
            step(it, false);
        }
    }
}

}  // namespace

__global__ void __launch_bounds__(kBjWarps * 32, BJ_MIN_CTAS) ot_join_blocks_kernel(const BjArgs a) {
    extern __shared__ __align__(128) unsigned char bj_smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    BjWarpSmem &ws = reinterpret_cast<BjWarpSmem *>(bj_smem)[warp];
    if (lane == 0) { mbar_init(&ws.bar[0], 1); mbar_init(&ws.bar[1], 1); }
    uint32_t qn = 0;  // results parked in ws.q (warp-uniform)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t parity = 0;  // bit s: phase of bar[s] the next wait on it looks for
    unsigned long long compared = 0;  // statistics: entries x slots compared (per lane, summed at the end)
    const uint32_t G = a.group;
    constexpr uint32_t kBlockBytes = kBjBlockWords * 4u;

    // first piece of a bucket: its blocks and its entries (16-byte aligned: buckets start at even entries)
    auto issue_first = [&](const uint4 m, int s) {
        if (lane == 0) {
            const uint32_t bytes_b = min(m.y, (uint32_t)kBjCapBlocks) * kBlockBytes;
            const uint32_t bytes_e = ((min(m.w, (uint32_t)kBjCapEnt) + 1u) >> 1) * 16u;
            mbar_expect_tx(&ws.bar[s], bytes_b + bytes_e);
            bulk_g2s(ws.st[s].blk, a.blocks + (size_t)m.x * kBjBlockWords, bytes_b, &ws.bar[s]);
            bulk_g2s(ws.st[s].ent, a.ent + m.z, bytes_e, &ws.bar[s]);
        }
    };

    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(a.ticket, 1u);
        t = __shfl_sync(kFull, t, 0);
        if (t >= a.n_groups) break;
        const uint32_t k0 = a.key_lo + t * G;
        const uint32_t nb = min(G, a.key_hi - k0);
        // bucket bounds of the group: lane i describes bucket k0 + i
        uint4 mine = make_uint4(0, 0, 0, 0);
        if (lane < nb) {
            const uint32_t b0 = __ldg(&a.blk_off[k0 + lane]) >> 6, b1 = __ldg(&a.blk_off[k0 + lane + 1]) >> 6;
            const uint32_t v0 = __ldg(&a.ent_off[k0 + lane]), v1 = __ldg(&a.ent_off[k0 + lane + 1]);
            mine = make_uint4(b0, b1 - b0, v0 & ~1u, (v1 & ~1u) - (v0 & ~1u) - (v0 & 1u));
            compared += (unsigned long long)mine.y * 64u * mine.w;
        }
        // buckets without entries or without windows have nothing to join
        uint32_t rest = __ballot_sync(kFull, mine.y != 0 && mine.w != 0);
        if (!rest) continue;
        if (lane < (uint32_t)kBjGroupMax) ws.meta[lane] = mine;
        __syncwarp();
        int s = 0;
        issue_first(ws.meta[__ffs((int)rest) - 1], 0);
        while (rest) {
            const uint4 m = ws.meta[__ffs((int)rest) - 1];
            rest &= rest - 1u;
            if (rest) issue_first(ws.meta[__ffs((int)rest) - 1], s ^ 1);
            const uint32_t K = m.w;
            for (uint32_t piece = 0; piece < m.y; piece += kBjCapBlocks) {
                const uint32_t nblk = min(m.y - piece, (uint32_t)kBjCapBlocks);
                if (piece) {  // a bucket larger than the stage: the further pieces are fetched in place
                    __syncwarp();
                    if (lane == 0) {
                        mbar_expect_tx(&ws.bar[s], nblk * kBlockBytes);
                        bulk_g2s(ws.st[s].blk, a.blocks + (size_t)(m.x + piece) * kBjBlockWords, nblk * kBlockBytes, &ws.bar[s]);
                    }
                }
                mbar_wait(&ws.bar[s], (parity >> s) & 1u);
                parity ^= 1u << s;
                for (uint32_t g0 = 0; g0 < K; g0 += 32) {
                    const uint32_t R = min(K - g0, 32u);
                    if (R > 16 || nblk == 1) bj_pass<1>(a, ws, ws.st[s], nblk, m.x + piece, m.z + g0, g0, R, qn);
                    else if (R > 8 || nblk == 2) bj_pass<2>(a, ws, ws.st[s], nblk, m.x + piece, m.z + g0, g0, R, qn);
                    else bj_pass<4>(a, ws, ws.st[s], nblk, m.x + piece, m.z + g0, g0, R, qn);
                }
            }
            __syncwarp();  // every lane is done with stage s before it is refilled
            s ^= 1;
        }
    }
    __syncwarp();
    if (qn) bj_flush(a, ws, qn, true);
    if (a.site_count) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) compared += __shfl_xor_sync(kFull, compared, o);
        if (lane == 0 && compared) atomicAdd(a.site_count + 1, compared);
    }
}

// ---- table side: slots -> bit-sliced blocks -------------------------------------------------------
//
// One warp per block of 64 slots: lane i reads slots i and 32 + i (pad slots carry gpos = 0xffffffff
// and differ from every letter, so they never hit), forty ballot pairs give the forty 64-bit words.
__global__ void __launch_bounds__(256) site_blocks_kernel(const uint4 *slots, uint32_t n_blocks, int dl, uint32_t *blocks) {
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t b = blockIdx.x * 8u + (threadIdx.x >> 5); b < n_blocks; b += gridDim.x * 8u) {
        const uint4 sa = slots[(size_t)b * 64u + lane], sb = slots[(size_t)b * 64u + 32u + lane];
        const bool pad_a = sa.x == 0xffffffffu, pad_b = sb.x == 0xffffffffu;
        uint2 w0 = make_uint2(0, 0), w1 = make_uint2(0, 0);  // lane l keeps words l and 32 + l
#pragma unroll
        for (int j = 0; j < 10; j++) {
            const uint32_t code_a = j < dl ? ((((sa.y >> j) & 1u) << 1) | ((sa.z >> j) & 1u)) : 0u;
            const uint32_t code_b = j < dl ? ((((sb.y >> j) & 1u) << 1) | ((sb.z >> j) & 1u)) : 0u;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const uint2 word = make_uint2(__ballot_sync(kFull, pad_a || code_a != (uint32_t)c),
                                              __ballot_sync(kFull, pad_b || code_b != (uint32_t)c));
                const int ix = 4 * j + c;
                if (ix < 32) { if (lane == (uint32_t)ix) w0 = word; }
                else if (lane == (uint32_t)(ix - 32)) w1 = word;
            }
        }
        uint2 *out = reinterpret_cast<uint2 *>(blocks + (size_t)b * kSiteBlockWords);
        out[lane] = w0;
        if (lane < kSiteBlockWords / 2 - 32) out[32 + lane] = w1;
    }
}

int launch_site_blocks(DeviceState *ds, const uint4 *d_slots, uint32_t n_blocks, int dl, uint32_t *d_blocks, cudaStream_t stream) {
    if (n_blocks == 0) return GUIDO_OK;
    const int grid = (int)std::min<uint32_t>((n_blocks + 7) / 8, (uint32_t)ds->sm_count * 16);
    site_blocks_kernel<<<grid, 256, 0, stream>>>(d_slots, n_blocks, dl, d_blocks);
    GUIDO_CUDA(cudaGetLastError());
    return GUIDO_OK;
}

int join_blocks_grid(DeviceState *ds, int *grid, int *smem) {
    *smem = (int)(kBjWarps * sizeof(BjWarpSmem));
    if (ds->bj_grid == 0) {
        GUIDO_CUDA(cudaFuncSetAttribute(ot_join_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, *smem));
        int per_sm = 0;
        GUIDO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ot_join_blocks_kernel, kBjWarps * 32, *smem));
        ds->bj_grid = ds->sm_count * std::max(per_sm, 1);
    }
    *grid = ds->bj_grid;
    return GUIDO_OK;
}

// joins the buckets [key_lo, key_hi); `ticket` must be zero when the kernel starts
int launch_join_blocks(DeviceState *ds, BjArgs a, cudaStream_t stream) {
    if (a.key_hi <= a.key_lo) return GUIDO_OK;
    int grid = 0, smem = 0, rc;
    if ((rc = join_blocks_grid(ds, &grid, &smem))) return rc;
    // groups small enough that every warp sees a few dozen tickets (the tail is then short), but at
    // least 4 buckets per ticket so the atomic and the bound loads stay off the critical path
    const uint32_t n_keys = a.key_hi - a.key_lo;
    const uint32_t warps = (uint32_t)grid * kBjWarps;
    uint32_t G = 16;
    while (G > 4 && n_keys / G < warps * 16u) G >>= 1;
    a.group = G;
    a.n_groups = (n_keys + G - 1) / G;
    ot_join_blocks_kernel<<<grid, kBjWarps * 32, smem, stream>>>(a);
    GUIDO_CUDA(cudaGetLastError());
    return GUIDO_OK;
}

}  // namespace guido
