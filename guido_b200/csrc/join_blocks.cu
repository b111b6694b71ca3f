// join_blocks.cu -- the off-target join for large guide sets: guide-side seed index (entries in
// bucket order) x genome-side PAM-site table in BIT-SLICED BLOCKS OF 32 WINDOWS.
//
// Sits under run_bowtie (reference guido/off_targets.py:89-227), like offtarget.cu.
//
// Both sides are grouped by core k-mer (offtarget.cu): a bucket of the site table holds the
// PAM-valid windows with that core, the same bucket of the seed index the guides whose core is
// within the seed budget of it.  What is left per bucket is W windows x K entries distal compares
// (human genome, 100 k guides: W ~ 366, K ~ 42).  The table is guide independent, so IT is the side
// that is bit-sliced, once, when it is built: a block of 32 windows stores, per distal base j and
// letter c, the word "window i differs from c at base j" (bit i) -- 40 words, 160 bytes, 5 bytes
// per window.  Buckets are padded to whole blocks (366 -> 384: 5 % pad slots instead of the 35 % of
// the round-1 layout that sliced the ~42 keys of a bucket into 64 slots).
//
// One LANE owns one index entry (a guide filed under this core): its ten letters select ten words
// of a block -- ten 4-byte shared-memory loads at per-lane addresses fixed for the whole bucket --
// and an 18-LOP3 carry-save tree says for which of the 32 windows (distal mismatches + bias)
// reaches 8, bias = core mismatches of the entry + 7 - total_max: no XOR/OR stage, no POPC, no
// per-window set-up.  Entries beyond a multiple of 32 do not get a whole pass of their own: 16 or
// fewer left-over entries share the lanes 2 or 4 ways and each copy takes every 2nd / 4th block.
//
// Data movement: a WARP streams its buckets by itself.  Lane 0 fetches the blocks of the next
// bucket with ONE bulk-TMA copy (cp.async.bulk -> UBLKCP, <= 2560 bytes, completion on the warp's
// own mbarrier) into the other half of a double buffer while the warp compares the current one; no
// CTA barrier anywhere in the loop.  Buckets are handed out in groups by an atomic ticket.
// Non-empty result words are queued per warp as (word, entry, block) and turned into hit records
// by all 32 lanes, 32 at a time (one atomic on the global cursor per flush).
#include "scan_common.cuh"
#include "join_blocks.cuh"

namespace guido {

namespace {

constexpr int kBjWarps = 8;                      // warps per CTA (they do not interact)
constexpr int kBjCapBlocks = 16;                 // blocks per stage buffer (larger buckets go in pieces)
constexpr int kBjBlockWords = kSiteBlockWords;   // 40
constexpr int kBjQueue = 96;                     // result words parked per warp (flushed from 64 on)

struct __align__(16) BjWarpSmem {
    uint32_t blk[2][kBjCapBlocks * kBjBlockWords];  // 2 x 2560 bytes, filled by bulk copies
    unsigned long long bar[2];                      // mbarriers: "buffer s is full"
    uint32_t q_r[kBjQueue], q_e[kBjQueue], q_b[kBjQueue];
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

__device__ __forceinline__ uint32_t lop_xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ uint32_t lop_maj(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ uint32_t lop_nor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0x01;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}

// what one lane keeps for the whole bucket: the ten word addresses of ITS letters in block 0 of the
// stage buffer and the four bias words (bit planes of the entry's bias, 0 or ~0)
struct LaneKey {
    const uint32_t *p[10];
    uint32_t q1, q2, q4, q8;
};

// bit i set <=> window i of the block `word_off` words further on is a hit for this lane's entry:
// carry-save count of the ten mismatch words plus the bias; only "reaches 8" matters, so sums that
// feed nothing are never formed (18 LOP3)
__device__ __forceinline__ uint32_t block_hits(const LaneKey &k, int word_off) {
    uint32_t m[10];
#pragma unroll
    for (int j = 0; j < 10; j++) m[j] = k.p[j][word_off + 4 * j];
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

// queued result words -> hit records.  Lane i takes queue entry i: the guide id comes from the
// index entry, position and planes of each hit window from its slot of the table.
__device__ __noinline__ void bj_flush(const BjArgs &a, BjWarpSmem &ws, uint32_t qn) {
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t b0 = 0; b0 < qn; b0 += 32) {
        const uint32_t i = b0 + lane;
        uint32_t r = i < qn ? ws.q_r[i] : 0u;
        const uint32_t e = i < qn ? ws.q_e[i] : 0u, gb = i < qn ? ws.q_b[i] : 0u;
        const uint32_t n = (uint32_t)__popc(r);
        uint32_t inc = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, inc, d);
            if (lane >= (uint32_t)d) inc += t;
        }
        const uint32_t total = __shfl_sync(kFull, inc, 31);
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.count, (unsigned long long)total);
        base = __shfl_sync(kFull, base, 0);
        unsigned long long dst = base + (inc - n);
        if (r) {
            const uint32_t guide = __ldg(&a.ent[e]).y;
            while (r) {
                const uint32_t bit = (uint32_t)__ffs((int)r) - 1u;
                r &= r - 1u;
                const uint4 s = __ldg(&a.slots[(size_t)gb * 32u + bit]);
                if (dst < a.cap) reinterpret_cast<uint4 *>(a.hits)[dst] = make_uint4(guide, s.x, s.y, s.z);
                dst++;
            }
        }
    }
    __syncwarp();
}

// one pass of <= 32 entries (R of them valid) over the nblk blocks of the stage buffer; with
// S = 2 / 4 the lanes hold S copies of <= 16 / 8 entries and copy s takes blocks s, s + S, ...
template <int S>
__device__ __forceinline__ void bj_pass(const BjArgs &a, BjWarpSmem &ws, const uint32_t *buf, uint32_t nblk,
                                        uint32_t gblk0, uint32_t e_first, uint32_t R, uint32_t &qn) {
    constexpr uint32_t kPer = 32 / S;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r_ix = lane & (kPer - 1), s_ix = lane / kPer;
    const bool valid = r_ix < R;
    const uint32_t e = e_first + (valid ? r_ix : 0u);
    const uint32_t w = valid ? __ldg(&a.ent[e]).x : (15u << 20);  // bias 15 never hits
    LaneKey k;
#pragma unroll
    for (int j = 0; j < 10; j++) k.p[j] = buf + s_ix * kBjBlockWords + ((w >> (2 * j)) & 3u);
    k.q1 = 0u - ((w >> 20) & 1u); k.q2 = 0u - ((w >> 21) & 1u); k.q4 = 0u - ((w >> 22) & 1u); k.q8 = 0u - ((w >> 23) & 1u);
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int it = 0; it < kBjCapBlocks / S; it++) {
        if ((uint32_t)(it * S) >= nblk) break;
        uint32_t r = block_hits(k, it * S * kBjBlockWords);
        if (S > 1 && it * S + s_ix >= nblk) r = 0;  // this copy's block is beyond the bucket
        const uint32_t any = __ballot_sync(kFull, r != 0);
        if (any) {
            if (r) {
                const uint32_t pos = qn + (uint32_t)__popc(any & lt);
                ws.q_r[pos] = r; ws.q_e[pos] = e; ws.q_b[pos] = gblk0 + it * S + s_ix;
            }
            qn += (uint32_t)__popc(any);
            if (qn > kBjQueue - 32) {
                __syncwarp();
                bj_flush(a, ws, qn);
                qn = 0;
            }
        }
    }
}

}  // namespace

__global__ void __launch_bounds__(kBjWarps * 32, 4) ot_join_blocks_kernel(const BjArgs a) {
    extern __shared__ __align__(128) unsigned char bj_smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    BjWarpSmem &ws = reinterpret_cast<BjWarpSmem *>(bj_smem)[warp];
    if (lane == 0) { mbar_init(&ws.bar[0], 1); mbar_init(&ws.bar[1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t parity = 0;  // bit s: phase of bar[s] the next wait on it looks for
    uint32_t qn = 0;      // queued result words (warp-uniform)
    unsigned long long compared = 0;  // statistics: entries x slots compared (lane 0)
    const uint32_t G = a.group;

    // a unit = one piece (<= kBjCapBlocks blocks) of one bucket
    auto issue = [&](uint32_t gblk, uint32_t nblk, int s) {
        if (lane == 0) {
            const uint32_t bytes = nblk * (kBjBlockWords * 4u);
            mbar_expect_tx(&ws.bar[s], bytes);
            bulk_g2s(ws.blk[s], a.blocks + (size_t)gblk * kBjBlockWords, bytes, &ws.bar[s]);
        }
    };

    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(a.ticket, 1u);
        t = __shfl_sync(kFull, t, 0);
        if (t >= a.n_groups) break;
        const uint32_t k0 = a.key_lo + t * G;
        const uint32_t nb = min(G, a.key_hi - k0);
        // bucket bounds of the group: lane i holds bucket k0 + i
        uint32_t b_lo = 0, b_hi = 0, e_lo = 0, e_hi = 0;
        if (lane < nb) {
            b_lo = __ldg(&a.blk_off[k0 + lane]) >> 5; b_hi = __ldg(&a.blk_off[k0 + lane + 1]) >> 5;
            e_lo = __ldg(&a.ent_off[k0 + lane]); e_hi = __ldg(&a.ent_off[k0 + lane + 1]);
        }
        // buckets without entries or without windows have nothing to join
        const uint32_t live = __ballot_sync(kFull, lane < nb && b_hi > b_lo && e_hi > e_lo);
        if (!live) continue;
        // unit cursor: bucket `ub` (lane index), blocks from `ublk`
        uint32_t rest = live;
        int ub = __ffs((int)rest) - 1;
        uint32_t cur_blk = __shfl_sync(kFull, b_lo, ub), cur_end = __shfl_sync(kFull, b_hi, ub);
        int s = 0;
        issue(cur_blk, min(cur_end - cur_blk, (uint32_t)kBjCapBlocks), 0);
        while (true) {
            const uint32_t nblk = min(cur_end - cur_blk, (uint32_t)kBjCapBlocks);
            // the unit after this one: next piece of the bucket, or the next live bucket of the group
            int nub = ub;
            uint32_t nxt_blk = cur_blk + nblk, nxt_end = cur_end, nrest = rest;
            if (nxt_blk >= cur_end) {
                nrest = rest & (rest - 1u);
                nub = nrest ? __ffs((int)nrest) - 1 : -1;
                if (nub >= 0) { nxt_blk = __shfl_sync(kFull, b_lo, nub); nxt_end = __shfl_sync(kFull, b_hi, nub); }
            }
            if (nub >= 0) issue(nxt_blk, min(nxt_end - nxt_blk, (uint32_t)kBjCapBlocks), s ^ 1);
            const uint32_t e0 = __shfl_sync(kFull, e_lo, ub), K = __shfl_sync(kFull, e_hi, ub) - e0;
            mbar_wait(&ws.bar[s], (parity >> s) & 1u);
            parity ^= 1u << s;
            const uint32_t *buf = ws.blk[s];
            if (lane == 0) compared += (unsigned long long)K * nblk * 32u;
            for (uint32_t g0 = 0; g0 < K; g0 += 32) {
                const uint32_t R = min(K - g0, 32u);
                if (R > 16 || nblk == 1) bj_pass<1>(a, ws, buf, nblk, cur_blk, e0 + g0, R, qn);
                else if (R > 8 || nblk == 2) bj_pass<2>(a, ws, buf, nblk, cur_blk, e0 + g0, R, qn);
                else bj_pass<4>(a, ws, buf, nblk, cur_blk, e0 + g0, R, qn);
            }
            __syncwarp();  // every lane is done with buffer s before it is refilled
            if (nub < 0) break;
            ub = nub; rest = nrest; cur_blk = nxt_blk; cur_end = nxt_end; s ^= 1;
        }
    }
    if (qn) bj_flush(a, ws, qn);
    if (a.site_count && lane == 0 && compared) atomicAdd(a.site_count + 1, compared);
}

// ---- table side: slots -> bit-sliced blocks -------------------------------------------------------
//
// One warp per block of 32 slots: lane i reads slot i (pad slots carry gpos = 0xffffffff and differ
// from every letter, so they never hit), forty ballots give the forty words.
__global__ void __launch_bounds__(256) site_blocks_kernel(const uint4 *slots, uint32_t n_blocks, int dl, uint32_t *blocks) {
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t b = blockIdx.x * 8u + (threadIdx.x >> 5); b < n_blocks; b += gridDim.x * 8u) {
        const uint4 s = slots[(size_t)b * 32u + lane];
        const bool pad = s.x == 0xffffffffu;
        uint32_t w0 = 0, w1 = 0;  // lane l keeps words l and 32 + l
#pragma unroll
        for (int j = 0; j < 10; j++) {
            const uint32_t code = j < dl ? ((((s.y >> j) & 1u) << 1) | ((s.z >> j) & 1u)) : 0u;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const uint32_t word = __ballot_sync(kFull, pad || code != (uint32_t)c);
                const int ix = 4 * j + c;
                if (ix < 32) { if (lane == (uint32_t)ix) w0 = word; }
                else if (lane == (uint32_t)(ix - 32)) w1 = word;
            }
        }
        uint32_t *out = blocks + (size_t)b * kSiteBlockWords;
        out[lane] = w0;
        if (lane < kSiteBlockWords - 32) out[32 + lane] = w1;
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
    uint32_t G = 32;
    while (G > 4 && n_keys / G < warps * 16u) G >>= 1;
    a.group = G;
    a.n_groups = (n_keys + G - 1) / G;
    ot_join_blocks_kernel<<<grid, kBjWarps * 32, smem, stream>>>(a);
    GUIDO_CUDA(cudaGetLastError());
    return GUIDO_OK;
}

}  // namespace guido
