// pam_scan.cu -- whole-genome / region PAM scan, both strands, ordered compaction.
//
// Sits under Locus._find_guides_in_interval (reference guido/locus.py:178-183): the two
// overlapping `re.finditer` passes become one streaming pass over the bit-plane image.
//
// One CTA = one tile of 256 blocks (16384 bases); a thread owns one 64-base block:
//   1. 128-bit coalesced load of {lo,hi}; the next block's words come from lane+1 by shuffle
//      (lane 31 re-loads, an L1 hit) -- that is the PAM-length halo;
//   2. per-strand 64-bit hit masks by bit-parallel set membership, AND-ed across PAM offsets;
//   3. run veto only in the rare tiles that touch a non-ACGT run / contig pad;
//   4. CTA prefix sum of popcounts, then a single-pass chained scan across tiles (decoupled
//      look-back on 64-bit {flag,count} descriptors; tiles are handed out by an atomic ticket so
//      a waiting tile's predecessors are always resident) => ascending output, one genome read;
//   5. hits staged in shared memory and written as coalesced 4-byte stores.
//
// HBM traffic per launch (algorithmic): 0.25 B/base read + 4 B per hit written.
#include <algorithm>

#include "scan_common.cuh"

namespace guido {

constexpr int kStage = 2048;  // staged hits per strand and tile (NGG: ~1024 expected)
constexpr unsigned long long kFlagAgg = 1ull << 62, kFlagPrefix = 2ull << 62, kValMask = (1ull << 62) - 1;

struct PamScanArgs {
    PlanesView pv;
    PamSpec pam;
    ScanRange r;
    int mode;
    int64_t first_block;  // global block index of tile 0's first block
    uint32_t n_tiles;
    uint32_t *out_fw, *out_rv;
    unsigned long long cap_fw, cap_rv;
    unsigned long long *counts;  // [0] n_fw, [1] n_rv
    unsigned long long *desc;    // 2 * n_tiles
    unsigned int *ticket;
};

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}

// chained-scan look-back for one strand, run by one full warp; returns the exclusive prefix
__device__ __forceinline__ unsigned long long lookback(unsigned long long *desc, uint32_t tile,
                                                       int strand, unsigned long long agg) {
    int lane = threadIdx.x & 31;
    unsigned long long *mine = desc + 2ull * tile + strand;
    if (tile == 0) {
        if (lane == 0) *reinterpret_cast<volatile unsigned long long *>(mine) = kFlagPrefix | agg;
        return 0;
    }
    if (lane == 0) *reinterpret_cast<volatile unsigned long long *>(mine) = kFlagAgg | agg;
    unsigned long long excl = 0;
    int64_t look = (int64_t)tile - 1;
    while (true) {
        int64_t idx = look - lane;
        unsigned long long d = kFlagPrefix;  // tiles before the first contribute a zero prefix
        if (idx >= 0) {
            const unsigned long long *p = desc + 2ull * (uint64_t)idx + strand;
            do { d = ld_volatile_u64(p); } while ((d >> 62) == 0);
        }
        unsigned pm = __ballot_sync(kFull, (d >> 62) == 2);
        int first = pm ? (__ffs(pm) - 1) : 32;
        unsigned long long v = (lane <= first) ? (d & kValMask) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        excl += v;
        if (pm) break;
        look -= 32;
    }
    if (lane == 0) *reinterpret_cast<volatile unsigned long long *>(mine) = kFlagPrefix | (excl + agg);
    return excl;
}

__global__ void __launch_bounds__(kThreads) pam_scan_kernel(const PamScanArgs a) {
    __shared__ uint32_t s_warp[kThreads / 32];
    __shared__ uint32_t s_tile;
    __shared__ int s_run[3];                     // run range [lo,hi) touching this tile, skip flag
    __shared__ unsigned long long s_excl[2];
    __shared__ uint32_t s_stage[2][kStage];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int P = a.pam.len;

    while (true) {
        if (threadIdx.x == 0) s_tile = atomicAdd(a.ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= a.n_tiles) break;
        const int64_t tile_block = a.first_block + (int64_t)tile * kThreads;
        const uint32_t t0 = (uint32_t)(tile_block * 64), t1 = t0 + kTileBases;
        if (threadIdx.x == 0) {
            int r_lo = first_run_ending_after(a.pv, t0 >= 32 ? t0 - 32 : 0);
            int r_hi = r_lo;
            int skip = 0;
            while (r_hi < a.pv.n_runs && a.pv.run_start[r_hi] < t1 + 32) {
                if (a.pv.run_start[r_hi] <= t0 && a.pv.run_end[r_hi] >= t1) skip = 1;
                r_hi++;
            }
            s_run[0] = r_lo; s_run[1] = r_hi; s_run[2] = skip;
        }
        __syncthreads();
        const int r_lo = s_run[0], r_hi = s_run[1];
        uint64_t hf = 0, hr = 0;
        const int64_t block = tile_block + threadIdx.x;
        if (!s_run[2]) {
            const ulonglong2 cur = load_block(a.pv, block);
            ulonglong2 nxt = shfl_down_u2(cur, 1);
            if (lane == 31) nxt = load_block(a.pv, block + 1);
            const uint64_t rm = range_mask(block, a.r, P);
            hf = pam_hit_mask(cur, nxt, a.pam.fw, P) & rm;
            hr = pam_hit_mask(cur, nxt, a.pam.rv, P) & rm;
            if (r_hi > r_lo && (hf | hr)) veto_hits(a.pv, r_lo, r_hi, a.mode, P, (uint32_t)(block * 64), hf, hr);
        }
        const uint32_t cf = __popcll(hf), cr = __popcll(hr);
        uint32_t total;
        const uint32_t excl = block_exclusive_scan(cf | (cr << 16), s_warp, &total);
        const uint32_t agg_f = total & 0xffff, agg_r = total >> 16;
        if (warp == 0) {
            unsigned long long e = lookback(a.desc, tile, 0, agg_f);
            if (lane == 0) s_excl[0] = e;
        } else if (warp == 1) {
            unsigned long long e = lookback(a.desc, tile, 1, agg_r);
            if (lane == 0) s_excl[1] = e;
        }
        // stage (or, for unusually dense tiles, write straight from registers)
        const uint32_t base = (uint32_t)(block * 64) - a.r.out_base;
        const bool stage_f = agg_f <= kStage, stage_r = agg_r <= kStage;
        uint32_t of = excl & 0xffff, orv = excl >> 16;
        if (stage_f) {
            uint64_t m = hf;
            while (m) { int i = __ffsll((long long)m) - 1; m &= m - 1; s_stage[0][of++] = base + i; }
        }
        if (stage_r) {
            uint64_t m = hr;
            while (m) { int i = __ffsll((long long)m) - 1; m &= m - 1; s_stage[1][orv++] = base + i; }
        }
        __syncthreads();
        const unsigned long long ef = s_excl[0], er = s_excl[1];
        if (stage_f) {
            for (uint32_t j = threadIdx.x; j < agg_f; j += kThreads)
                if (ef + j < a.cap_fw) a.out_fw[ef + j] = s_stage[0][j];
        } else {
            uint64_t m = hf;
            unsigned long long o = ef + of;
            while (m) { int i = __ffsll((long long)m) - 1; m &= m - 1; if (o < a.cap_fw) a.out_fw[o] = base + i; o++; }
        }
        if (stage_r) {
            for (uint32_t j = threadIdx.x; j < agg_r; j += kThreads)
                if (er + j < a.cap_rv) a.out_rv[er + j] = s_stage[1][j];
        } else {
            uint64_t m = hr;
            unsigned long long o = er + orv;
            while (m) { int i = __ffsll((long long)m) - 1; m &= m - 1; if (o < a.cap_rv) a.out_rv[o] = base + i; o++; }
        }
        if (tile == a.n_tiles - 1 && threadIdx.x == 0) {
            a.counts[0] = ef + agg_f;
            a.counts[1] = er + agg_r;
        }
        __syncthreads();  // s_stage / s_tile are rewritten by the next iteration
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
    a.pv = pv; a.pam = pam; a.r = r; a.mode = mode;
    a.first_block = r.lo / 64;
    int64_t last_block = ((int64_t)r.hi - 1) / 64;
    a.n_tiles = (uint32_t)((last_block - a.first_block + 1 + kThreads - 1) / kThreads);
    a.out_fw = d_fw; a.out_rv = d_rv; a.cap_fw = cap_fw; a.cap_rv = cap_rv;
    a.counts = d_counts;
    int64_t need = 2 * (int64_t)a.n_tiles;
    if (need > ds->desc_cap) {
        if (ds->d_desc) cudaFree(ds->d_desc);
        ds->d_desc = nullptr; ds->desc_cap = 0;
        GUIDO_CUDA(cudaMalloc(&ds->d_desc, (size_t)need * sizeof(unsigned long long)));
        ds->desc_cap = need;
    }
    a.desc = ds->d_desc;
    a.ticket = ds->d_tile_counter;
    GUIDO_CUDA(cudaMemsetAsync(ds->d_desc, 0, (size_t)need * sizeof(unsigned long long), stream));
    GUIDO_CUDA(cudaMemsetAsync(ds->d_tile_counter, 0, sizeof(unsigned int), stream));
    int per_sm = 0;
    GUIDO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pam_scan_kernel, kThreads, 0));
    int grid = ds->sm_count * std::max(per_sm, 1);
    if ((uint32_t)grid > a.n_tiles) grid = (int)a.n_tiles;
    pam_scan_kernel<<<grid, kThreads, 0, stream>>>(a);
    GUIDO_CUDA(cudaGetLastError());
    if (n_launches) *n_launches += 1;
    return GUIDO_OK;
}

}  // namespace guido
