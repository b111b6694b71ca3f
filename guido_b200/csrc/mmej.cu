// mmej.cu -- microhomology (MMEJ) candidate enumeration, one warp per cut site.
//
// Sits under generate_mmej_patterns (reference guido/mmej.py:8-64).  The reference tries every
// k-mer of the left flank (k = len(left)-1 .. 2), keeps it when it also occurs in the right flank
// and was not seen before, lists its greedy non-overlapping occurrences on both sides
// (re.finditer), takes their product, and finally drops every candidate that an EARLIER candidate
// contains on both sides.  The same result, in the same order, without string objects:
//
//   * rmax[i]  = longest prefix of left[i:] that occurs anywhere in right
//     lprev[i] = longest prefix of left[i:] that also starts at some i' < i in left
//     => k-mer (i,k) is accepted  <=>  lprev[i] < k <= rmax[i]   (first occurrence, present right)
//   * for k descending, i ascending: occurrences by lane-parallel compare + warp ballot, greedy
//     non-overlap selection on the ballot words, product emitted left-major;
//   * elimination: lane j scans candidates 0..j-1 (broadcast reads), exactly the reference's
//     upper-triangular containment test (mmej.py:54-64).
//
// Scores are float64 functions of small integers and stay on the host (bit-exact with Python).
#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "device.cuh"

namespace guido {

constexpr int kMmejWarps = 4;        // warps per CTA
constexpr int kMaxSide = 255;        // max letters on either side of the cut
constexpr unsigned kFullMask = 0xffffffffu;

struct MmejArgs {
    const char *windows;
    const long long *offsets;   // n+1
    const int *cuts;            // n
    int n;
    uint32_t *scratch;          // per warp: cap candidates + cap/32 keep words
    uint32_t cap;               // candidates per warp
    uint32_t *out;              // kept tuples, segments in completion order
    unsigned long long out_cap;
    unsigned long long *cursor;
    long long *seg_start;       // n: start of guide g's segment in out (or -1 on scratch overflow)
    int *seg_count;             // n: kept tuples
    int *n_cand;                // n: candidates before elimination
    const int *todo;            // optional list of window indices to process (retry pass)
    int n_todo;
};

struct WarpSmem {
    unsigned char left[256], right[256];
    unsigned char rmax[256], lprev[256];
    unsigned char lpos[128], rpos[128];
    uint32_t ml[8], mr[8];
};

// greedy left-to-right non-overlapping starts out of a match bitmask (re.finditer semantics)
__device__ __forceinline__ int greedy_select(const uint32_t *mask, int n_words, int n_pos, int k, unsigned char *out, int lane) {
    int count = 0, next = 0;
    while (next < n_pos) {
        int w = next >> 5;
        uint32_t m = mask[w] & (kFullMask << (next & 31));
        while (!m && ++w < n_words) m = mask[w];
        if (!m || w >= n_words) break;
        int x = (w << 5) + __ffs(m) - 1;
        if (x >= n_pos) break;
        if (lane == 0) out[count] = (unsigned char)x;
        count++;
        next = x + k;
    }
    return count;
}

__global__ void __launch_bounds__(kMmejWarps * 32) mmej_kernel(const MmejArgs a) {
    __shared__ WarpSmem s_all[kMmejWarps];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    WarpSmem &s = s_all[wib];
    const int warp_global = blockIdx.x * kMmejWarps + wib, n_warps = gridDim.x * kMmejWarps;
    uint32_t *cand = a.scratch + (size_t)warp_global * (a.cap + a.cap / 32);
    uint32_t *keep = cand + a.cap;
    const int n_items = a.todo ? a.n_todo : a.n;

    for (int it = warp_global; it < n_items; it += n_warps) {
        const int g = a.todo ? a.todo[it] : it;
        const long long off = a.offsets[g];
        const int len = (int)(a.offsets[g + 1] - off);
        int cut = a.cuts[g];
        cut = cut < 0 ? 0 : (cut > len ? len : cut);
        const int nl = cut, nr = len - cut;
        __syncwarp();
        for (int i = lane; i < nl; i += 32) s.left[i] = (unsigned char)a.windows[off + i];
        for (int i = lane; i < nr; i += 32) s.right[i] = (unsigned char)a.windows[off + cut + i];
        __syncwarp();
        // rmax / lprev
        int kmax = 0;
        for (int i = lane; i < nl; i += 32) {
            int best = 0;
            for (int j = 0; j < nr; j++) {
                int t = 0;
                while (i + t < nl && j + t < nr && s.left[i + t] == s.right[j + t]) t++;
                best = max(best, t);
            }
            int prev = 0;
            for (int j = 0; j < i; j++) {
                int t = 0;
                while (i + t < nl && s.left[i + t] == s.left[j + t]) t++;
                prev = max(prev, t);
            }
            s.rmax[i] = (unsigned char)best;
            s.lprev[i] = (unsigned char)prev;
            kmax = max(kmax, best);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) kmax = max(kmax, __shfl_xor_sync(kFullMask, kmax, o));
        kmax = min(kmax, nl - 1);
        __syncwarp();

        uint32_t nc = 0;
        for (int k = kmax; k >= 2; k--) {                       // mmej.py:24
            for (int ibase = 0; ibase + k <= nl; ibase += 32) {  // mmej.py:27
                const int i = ibase + lane;
                const bool acc = (i + k <= nl) && s.lprev[i] < k && s.rmax[i] >= k;
                uint32_t am = __ballot_sync(kFullMask, acc);
                while (am) {
                    const int i0 = ibase + __ffs(am) - 1;
                    am &= am - 1;
                    // occurrences of left[i0:i0+k] on both sides (only the mask words either side can fill)
                    const int n_words = (max(nl, nr) - k + 32) >> 5;
                    for (int w = 0; w < n_words; w++) {
                        int x = (w << 5) + lane;
                        bool ml = false, mr = false;
                        if (x + k <= nl) {
                            int t = 0;
                            while (t < k && s.left[x + t] == s.left[i0 + t]) t++;
                            ml = (t == k);
                        }
                        if (x + k <= nr) {
                            int t = 0;
                            while (t < k && s.right[x + t] == s.left[i0 + t]) t++;
                            mr = (t == k);
                        }
                        uint32_t bl = __ballot_sync(kFullMask, ml), br = __ballot_sync(kFullMask, mr);
                        if (lane == 0) { s.ml[w] = bl; s.mr[w] = br; }
                    }
                    __syncwarp();
                    const int cl = greedy_select(s.ml, n_words, nl - k + 1, k, s.lpos, lane);
                    const int cr = greedy_select(s.mr, n_words, nr - k + 1, k, s.rpos, lane);
                    __syncwarp();
                    const int np = cl * cr;                       // itertools.product, left-major
                    for (int t = lane; t < np; t += 32) {
                        uint32_t idx = nc + t;
                        if (idx < a.cap) {
                            uint32_t ls = s.lpos[t / cr], rs = s.rpos[t % cr];
                            cand[idx] = ls | ((ls + k) << 8) | (rs << 16) | ((rs + k) << 24);
                        }
                    }
                    nc += np;
                    __syncwarp();
                }
            }
        }
        if (lane == 0) a.n_cand[g] = (int)nc;
        if (nc > a.cap) {  // scratch too small for this (very repetitive) window: host retries it
            if (lane == 0) { a.seg_start[g] = -1; a.seg_count[g] = 0; }
            continue;
        }
        __syncwarp();
        // nested-pattern elimination (mmej.py:54-64)
        uint32_t kept_total = 0;
        for (uint32_t jbase = 0; jbase < nc; jbase += 32) {
            const uint32_t j = jbase + lane;
            bool alive = j < nc;
            uint32_t cj = alive ? cand[j] : 0;
            const uint32_t ls = cj & 255, le = (cj >> 8) & 255, rs = (cj >> 16) & 255, re = cj >> 24;
            const uint32_t jmax = min(jbase + 32, nc);
            for (uint32_t i = 0; i + 1 < jmax; i++) {
                const uint32_t ci = cand[i];
                if (alive && i < j && (ci & 255) <= ls && ((ci >> 8) & 255) >= le &&
                    ((ci >> 16) & 255) <= rs && (ci >> 24) >= re)
                    alive = false;
                if (!__any_sync(kFullMask, alive)) break;
            }
            const uint32_t km = __ballot_sync(kFullMask, alive);
            if (lane == 0) keep[jbase >> 5] = km;
            kept_total += __popc(km);
        }
        unsigned long long base = 0;
        if (lane == 0) {
            base = atomicAdd(a.cursor, (unsigned long long)kept_total);
            a.seg_start[g] = (long long)base;
            a.seg_count[g] = (int)kept_total;
        }
        base = __shfl_sync(kFullMask, base, 0);
        __syncwarp();
        uint32_t written = 0;
        for (uint32_t jbase = 0; jbase < nc; jbase += 32) {
            const uint32_t km = keep[jbase >> 5];
            if ((km >> lane) & 1) {
                unsigned long long o = base + written + __popc(km & ((1u << lane) - 1));
                if (o < a.out_cap) a.out[o] = cand[jbase + lane];
            }
            written += __popc(km);
        }
    }
}


// ---- second draft: bit-parallel enumeration (mmej_fast_kernel) ------------------------------------
//
// Same result, same order as mmej_kernel above, for windows of A/C/G/T with at most 127 letters on
// either side of the cut (everything guido produces with its default 75-base flanks); any other
// window, and any window that overflows one of the fixed-size work areas, is flagged
// (seg_start = -2) and goes through the general kernel.
//
//   * per letter, the positions where it occurs on each side as a 128-bit mask (8 ballots a side);
//   * rmax / lprev by Shift-And: lane i keeps the 128-bit sets "starts j where right[j..j+t) ==
//     left[i..i+t)" and "starts j < i where left[j..j+t) == left[i..i+t)" and ANDs in one shifted
//     letter mask per step -- ~5 steps for random DNA instead of 75 x 75 byte loops;
//   * the accepted (i, k) pairs are compacted into a work list in the reference's order (k
//     descending, i ascending); then ONE LANE PER K-MER: occurrence masks by k Shift-And steps,
//     greedy non-overlapping selection in registers (re.finditer), positions parked in per-lane
//     byte lists; the products are written by all 32 lanes, output slot -> (k-mer, left
//     occurrence, right occurrence) by a 5-step search over the round's prefix sums;
//   * elimination: a candidate can only be contained in a LONGER one (equal length would mean the
//     same tuple), so class k is tested against the candidates before the class only, four per
//     128-bit load; with tuples stored as bytes (ls, le, rs, re) <= 127 containment is one
//     subtraction and one mask: t = tuple ^ 0x7f007f00 turns "le' >= le" into "<=" like the
//     starts, and ((t | 0x80808080) - t') keeps all four guard bits iff every byte of t is >= t'.
constexpr int kFastSide = 127;       // letters per side
constexpr int kFastWords = 4;        // 128-bit position sets
constexpr int kFastWork = 1024;      // accepted k-mers per window
constexpr int kFastList = 40;        // occurrences per side of one k-mer kept in the per-lane lists
constexpr int kFastParents = 512;    // leading candidates mirrored in shared memory for the elimination
constexpr uint32_t kTupleFlip = 0x7f007f00u, kTupleGuard = 0x80808080u;

struct FastArgs {
    MmejArgs m;
    // windows read from a resident genome image instead of letters (m.windows == nullptr):
    // window g = global positions [starts[g], starts[g] + lens[g]), m.cuts[g] relative to its start
    PlanesView pv;
    const uint32_t *starts;
    const int *lens;
};

struct __align__(16) FastSmem {
    uint32_t lm[4][2 * kFastWords], rm[4][2 * kFastWords];  // letter masks, upper halves zero (shifted reads stay inside)
    uint8_t left[128], right[128];                         // 2-bit codes
    uint8_t rmax[128], lprev[128];
    uint16_t work[kFastWork];                              // i | k << 8
    uint8_t lpos[kFastList][32], rpos[kFastList][32];
    uint32_t parents[kFastParents];                        // tuple ^ kTupleFlip
    uint32_t class_n[128];                                 // candidates of length k, then: candidates longer than k
    uint32_t pfx[33];                                      // exclusive prefix of the round's products
    uint32_t lane_k[32], lane_cr[32], lane_rc[32];         // per k-mer of the round: k, right occurrences, 65536 / cr + 1
};

// word w of (128-bit mask of letter c) >> t
__device__ __forceinline__ uint32_t mask_shr(const uint32_t (*m)[2 * kFastWords], uint32_t c, int w, int t) {
    const uint32_t *row = m[c] + w + (t >> 5);
    return __funnelshift_r(row[0], row[1], t);  // shift by t & 31
}

// greedy left-to-right non-overlapping starts of a 128-bit occurrence set (re.finditer); positions
// beyond `cap` are counted, not stored
__device__ __forceinline__ int greedy128(const uint32_t (&o)[kFastWords], int k, uint8_t (*pos)[32], int lane, int cap) {
    int count = 0, next = 0;
#pragma unroll
    for (int w = 0; w < kFastWords; w++) {
        uint32_t m = o[w];
        const int lo = next - 32 * w;
        if (lo >= 32) m = 0;
        else if (lo > 0) m &= 0xffffffffu << lo;
        while (m) {
            const int b = __ffs((int)m) - 1;
            if (count < cap) pos[count][lane] = (uint8_t)(32 * w + b);
            count++;
            next = 32 * w + b + k;
            m = b + k >= 32 ? 0u : (m & (0xffffffffu << (b + k)));
        }
    }
    return count;
}

__global__ void __launch_bounds__(kMmejWarps * 32) mmej_fast_kernel(const FastArgs fa) {
    __shared__ FastSmem s_all[kMmejWarps];
    const MmejArgs &a = fa.m;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    FastSmem &s = s_all[wib];
    const int warp_global = blockIdx.x * kMmejWarps + wib, n_warps = gridDim.x * kMmejWarps;
    uint32_t *cand = a.scratch + (size_t)warp_global * (a.cap + a.cap / 32);
    uint32_t *keep = cand + a.cap;
    for (int w = lane; w < 4 * 2 * kFastWords; w += 32) { (&s.lm[0][0])[w] = 0; (&s.rm[0][0])[w] = 0; }

    for (int g = warp_global; g < a.n; g += n_warps) {
        __syncwarp();
        int len, cut;
        long long off = 0;
        uint32_t gstart = 0;
        if (a.windows) { off = a.offsets[g]; len = (int)(a.offsets[g + 1] - off); }
        else { gstart = fa.starts[g]; len = fa.lens[g]; }
        cut = a.cuts[g];
        cut = cut < 0 ? 0 : (cut > len ? len : cut);
        const int nl = cut, nr = len - cut;
        bool slow = nl > kFastSide || nr > kFastSide;
        if (!slow && !a.windows) {
            // a window that touches a non-ACGT run of the image has letters the planes do not hold
            int lo = 0, hi = fa.pv.n_runs;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(&fa.pv.run_end[mid]) <= gstart) lo = mid + 1; else hi = mid; }
            slow = lo < fa.pv.n_runs && __ldg(&fa.pv.run_start[lo]) < gstart + (uint32_t)len;
        }
        if (!slow) {
            bool bad = false;
            for (int i = lane; i < len; i += 32) {
                uint32_t code;
                if (a.windows) {
                    const char ch = a.windows[off + i];
                    code = ch == 'A' ? 0u : ch == 'C' ? 1u : ch == 'G' ? 2u : ch == 'T' ? 3u : 255u;
                } else {
                    const uint32_t x = gstart + (uint32_t)i;
                    const ulonglong2 b = fa.pv.planes[(int64_t)(x >> 6) - fa.pv.blk_lo];
                    code = (uint32_t)((b.x >> (x & 63)) & 1ull) | ((uint32_t)((b.y >> (x & 63)) & 1ull) << 1);
                }
                bad |= code == 255u;
                if (i < nl) s.left[i] = (uint8_t)code; else s.right[i - nl] = (uint8_t)code;
            }
            slow = __any_sync(kFullMask, bad);
        }
        if (slow) {
            if (lane == 0) { a.seg_start[g] = -2; a.seg_count[g] = 0; a.n_cand[g] = 0; }
            continue;
        }
        __syncwarp();
        // letter masks
#pragma unroll
        for (int w = 0; w < kFastWords; w++) {
            const int x = 32 * w + lane;
            const uint32_t cl = x < nl ? s.left[x] : 255u, cr = x < nr ? s.right[x] : 255u;
#pragma unroll
            for (uint32_t c = 0; c < 4; c++) {
                const uint32_t bl = __ballot_sync(kFullMask, cl == c), br = __ballot_sync(kFullMask, cr == c);
                if (lane == 0) { s.lm[c][w] = bl; s.rm[c][w] = br; }
            }
        }
        for (int k = lane; k < 128; k += 32) s.class_n[k] = 0;
        __syncwarp();
        // rmax / lprev by Shift-And, lane = left start
        int kmax = 0;
        for (int ibase = 0; ibase < nl; ibase += 32) {
            const int i = ibase + lane;
            uint32_t dr[kFastWords], dl[kFastWords];
#pragma unroll
            for (int w = 0; w < kFastWords; w++) {
                dr[w] = i < nl ? 0xffffffffu : 0u;
                const int below = i - 32 * w;  // starts j < i
                dl[w] = i >= nl || below <= 0 ? 0u : below >= 32 ? 0xffffffffu : ((1u << below) - 1u);
            }
            int rmx = 0, lpv = 0;
            for (int t = 0;; t++) {
                const bool in = i + t < nl;
                const uint32_t c = in ? s.left[i + t] : 0u;
                uint32_t any_r = 0, any_l = 0;
#pragma unroll
                for (int w = 0; w < kFastWords; w++) {
                    dr[w] = in ? dr[w] & mask_shr(s.rm, c, w, t) : 0u;
                    dl[w] = in ? dl[w] & mask_shr(s.lm, c, w, t) : 0u;
                    any_r |= dr[w]; any_l |= dl[w];
                }
                if (any_r) rmx = t + 1;
                if (any_l) lpv = t + 1;
                if (!__any_sync(kFullMask, (any_r | any_l) != 0u)) break;
            }
            if (i < nl) { s.rmax[i] = (uint8_t)rmx; s.lprev[i] = (uint8_t)lpv; }
            kmax = max(kmax, rmx);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) kmax = max(kmax, __shfl_xor_sync(kFullMask, kmax, o));
        kmax = min(kmax, nl - 1);
        __syncwarp();
        // work list: accepted (i, k), k descending, i ascending (mmej.py:24-32)
        uint32_t nw = 0;
        for (int k = kmax; k >= 2; k--) {
            for (int ibase = 0; ibase + k <= nl; ibase += 32) {
                const int i = ibase + lane;
                const bool acc = i + k <= nl && s.lprev[i] < k && s.rmax[i] >= k;
                const uint32_t am = __ballot_sync(kFullMask, acc);
                const uint32_t at = nw + (uint32_t)__popc(am & ((1u << lane) - 1u));
                if (acc && at < (uint32_t)kFastWork) s.work[at] = (uint16_t)(i | (k << 8));
                nw += (uint32_t)__popc(am);
            }
        }
        bool overflow = nw > (uint32_t)kFastWork;
        __syncwarp();
        // one lane per k-mer
        uint32_t nc = 0;
        for (uint32_t base = 0; base < nw && !overflow; base += 32) {
            const bool on = base + lane < nw;
            const uint32_t e = on ? s.work[base + lane] : 0u;
            const int i = (int)(e & 255u), k = (int)(e >> 8);
            const int k_round = (int)(s.work[base] >> 8);  // the round's longest k-mer
            uint32_t ol[kFastWords], orr[kFastWords];
#pragma unroll
            for (int w = 0; w < kFastWords; w++) ol[w] = orr[w] = on ? 0xffffffffu : 0u;
            for (int t = 0; t < k_round; t++) {
                if (t < k) {
                    const uint32_t c = s.left[i + t];
#pragma unroll
                    for (int w = 0; w < kFastWords; w++) { ol[w] &= mask_shr(s.lm, c, w, t); orr[w] &= mask_shr(s.rm, c, w, t); }
                }
            }
            const int cl = on ? greedy128(ol, k, s.lpos, lane, kFastList) : 0;
            const int cr = on ? greedy128(orr, k, s.rpos, lane, kFastList) : 0;
            overflow = __any_sync(kFullMask, cl > kFastList || cr > kFastList);
            if (overflow) break;
            const uint32_t np = (uint32_t)(cl * cr);
            uint32_t inc = np;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(kFullMask, inc, d); if (lane >= d) inc += u; }
            const uint32_t total = __shfl_sync(kFullMask, inc, 31);
            s.pfx[lane] = inc - np;
            s.lane_k[lane] = (uint32_t)k; s.lane_cr[lane] = (uint32_t)cr; s.lane_rc[lane] = cr ? 65536u / (uint32_t)cr + 1u : 0u;
            if (lane == 0) s.pfx[32] = total;
            if (np) atomicAdd(&s.class_n[k], np);
            __syncwarp();
            // products, left-major (itertools.product), written by all lanes
            for (uint32_t o = lane; o < total; o += 32) {
                uint32_t src = 0;
#pragma unroll
                for (uint32_t step = 16; step > 0; step >>= 1) if (s.pfx[src + step] <= o) src += step;
                const uint32_t r = o - s.pfx[src], crs = s.lane_cr[src], ks = s.lane_k[src];
                const uint32_t li = (r * s.lane_rc[src]) >> 16, ri = r - li * crs;  // r / cr, exact for r < 1600 (40 x 40)
                const uint32_t ls = s.lpos[li][src], rs = s.rpos[ri][src];
                const uint32_t tup = ls | ((ls + ks) << 8) | (rs << 16) | ((rs + ks) << 24);
                const uint32_t idx = nc + o;
                if (idx < a.cap) cand[idx] = tup;
                if (idx < (uint32_t)kFastParents) s.parents[idx] = tup ^ kTupleFlip;
            }
            nc += total;
            __syncwarp();
        }
        if (overflow) {
            if (lane == 0) { a.seg_start[g] = -2; a.seg_count[g] = 0; a.n_cand[g] = 0; }
            continue;
        }
        if (lane == 0) a.n_cand[g] = (int)nc;
        if (nc > a.cap) {  // scratch too small for this (very repetitive) window: the host retries it
            if (lane == 0) { a.seg_start[g] = -1; a.seg_count[g] = 0; }
            continue;
        }
        __syncwarp();
        // nested-pattern elimination (mmej.py:54-64), class by class: class k = candidates of length k,
        // in list order behind all longer ones
        uint32_t kept_total = 0;
        {
            uint32_t first = 0;  // candidates longer than k = start of class k
            for (int k = kmax; k >= 2; k--) {
                const uint32_t n_k = s.class_n[k];
                for (uint32_t jbase = first; jbase < first + n_k; jbase += 32) {
                    const uint32_t j = jbase + lane;
                    bool alive = j < first + n_k;
                    const uint32_t tg = ((alive ? cand[j] : 0u) ^ kTupleFlip) | kTupleGuard;
                    const uint32_t in_smem = min(first, (uint32_t)kFastParents);
                    // y = ~(tg - parent) & guard is zero iff the parent contains the candidate: four parents
                    // per 128-bit load, one unsigned minimum tree and one compare per group; the vote that
                    // ends the scan early (every lane already eliminated) runs every fourth group
                    const uint32_t full = in_smem & ~3u;
                    uint32_t i = 0;
                    for (; i < full; i += 4) {
                        const uint4 p = *reinterpret_cast<const uint4 *>(&s.parents[i]);
                        const uint32_t y = min(min(~(tg - p.x) & kTupleGuard, ~(tg - p.y) & kTupleGuard),
                                               min(~(tg - p.z) & kTupleGuard, ~(tg - p.w) & kTupleGuard));
                        alive = alive && y != 0u;
                        if ((i & 12u) == 12u && !__any_sync(kFullMask, alive)) { i = in_smem; break; }
                    }
                    for (; i < in_smem; i++)  // the last, partial group
                        alive = alive && (~(tg - s.parents[i]) & kTupleGuard) != 0u;
                    for (uint32_t i = in_smem; i < first; i++) {
                        const uint32_t pp = cand[i] ^ kTupleFlip;
                        alive = alive && ((tg - pp) & kTupleGuard) != kTupleGuard;
                    }
                    // class starts are not multiples of 32: the keep words are addressed per candidate
                    const uint32_t km = __ballot_sync(kFullMask, alive);
                    const uint32_t lo_bit = jbase & 31u, wi = jbase >> 5;
                    if (lane == 0) {
                        keep[wi] = (lo_bit ? keep[wi] : 0u) | (km << lo_bit);
                        if (lo_bit && wi + 1 < a.cap / 32) keep[wi + 1] = km >> (32u - lo_bit);
                    }
                    kept_total += (uint32_t)__popc(km);
                    __syncwarp();
                }
                first += n_k;
            }
        }
        unsigned long long base_out = 0;
        if (lane == 0) {
            base_out = atomicAdd(a.cursor, (unsigned long long)kept_total);
            a.seg_start[g] = (long long)base_out;
            a.seg_count[g] = (int)kept_total;
        }
        base_out = __shfl_sync(kFullMask, base_out, 0);
        __syncwarp();
        uint32_t written = 0;
        for (uint32_t jbase = 0; jbase < nc; jbase += 32) {
            const uint32_t km = keep[jbase >> 5];
            if ((km >> lane) & 1) {
                const unsigned long long o = base_out + written + __popc(km & ((1u << lane) - 1));
                if (o < a.out_cap) a.out[o] = cand[jbase + lane];
            }
            written += __popc(km);
        }
    }
}


// segments in completion order -> window order: counts copied for the scan, then one warp per window
__global__ void __launch_bounds__(256) mmej_counts_kernel(const int *seg_count, int n, uint32_t *counts) {
    const int g = blockIdx.x * 256 + threadIdx.x;
    if (g <= n) counts[g] = g < n ? (uint32_t)seg_count[g] : 0u;
}

__global__ void __launch_bounds__(256) mmej_order_kernel(const uint32_t *src, const long long *seg_start, const uint32_t *offs, int n,
                                                         uint32_t *dst, unsigned long long dst_cap) {
    const int lane = threadIdx.x & 31;
    for (int g = blockIdx.x * 8 + (threadIdx.x >> 5); g < n; g += gridDim.x * 8) {
        const long long from = seg_start[g];
        const uint32_t to = offs[g], cnt = offs[g + 1] - to;
        if (from < 0 || (unsigned long long)to + cnt > dst_cap) continue;
        for (uint32_t i = lane; i < cnt; i += 32) dst[to + i] = src[from + i];
    }
}

}  // namespace guido

using namespace guido;

namespace {

// Stream, events and device buffers are kept per device between calls (grown on demand): a
// locus-sized batch runs for milliseconds, a context and nine allocations per call cost as much
struct MmejCtx {
    DeviceState ds;
    void *win = nullptr, *off = nullptr, *cuts = nullptr, *seg_start = nullptr, *seg_count = nullptr, *ncand = nullptr,
         *todo = nullptr, *scratch = nullptr, *out = nullptr, *starts = nullptr, *lens = nullptr, *offs = nullptr, *fin = nullptr;
    int64_t win_cap = 0, off_cap = 0, cuts_cap = 0, seg_start_cap = 0, seg_count_cap = 0, ncand_cap = 0, todo_cap = 0,
            scratch_cap = 0, out_cap = 0, starts_cap = 0, lens_cap = 0, offs_cap = 0, fin_cap = 0;
};
std::mutex g_mmej_mu;
MmejCtx *g_mmej_ctx[64] = {};

int mmej_ctx(int device, MmejCtx **out) {
    if (device < 0 || device >= 64) { set_error("device %d out of range", device); return GUIDO_ERR_ARG; }
    if (!g_mmej_ctx[device]) {
        MmejCtx *c = new MmejCtx();
        int rc = device_init(&c->ds, device);
        if (rc) { delete c; return rc; }
        g_mmej_ctx[device] = c;
    }
    *out = g_mmej_ctx[device];
    return GUIDO_OK;
}

#define MM_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_error("CUDA error %s at %s:%d", cudaGetErrorName(e_), __FILE__, __LINE__); return GUIDO_ERR_CUDA; } } while (0)
#define MM_BYTES(p, need) do { if ((rc = ensure_bytes(&cx.p, &cx.p##_cap, (int64_t)(need)))) return rc; } while (0)

// Runs the enumeration for n windows given as letters (windows / offsets) or as stretches of a
// resident image (img / starts / lens).  Pass 0 is the bit-parallel kernel; whatever it hands back
// (seg_start -2: a letter outside ACGT, a side longer than 127, a work area overflow) goes through
// the general kernel, which needs letters -- image windows are fetched from the host copy of the
// image for that; windows whose candidates overflow the scratch (seg_start -1) are retried with
// 16x the scratch, as before.
int mmej_run(MmejCtx &cx, const char *windows, const int64_t *offsets, const guido_image *img, const uint32_t *starts,
             const int32_t *lens, const int32_t *cuts, int64_t n, uint32_t *out_tuples, int64_t cap, int64_t *out_offsets,
             int32_t *n_cand, guido_stats *stats) {
    DeviceState &ds = cx.ds;
    int rc;
    MM_CUDA(cudaSetDevice(ds.device));
    const bool from_image = img != nullptr;
    const int64_t total_len = from_image ? 0 : offsets[n];
    unsigned long long *d_cursor = ds.d_counts;
    std::vector<long long> seg_start((size_t)n);
    std::vector<int> seg_count((size_t)n), ncand((size_t)n), final_count((size_t)n, 0), final_ncand((size_t)n, 0);
    int launches = 0;
    float kernel_ms = 0;
    cudaStream_t st = ds.stream;
    MM_CUDA(cudaEventRecord(ds.ev[0], st));
    MM_BYTES(cuts, n * 4);
    MM_BYTES(seg_start, n * 8);
    MM_BYTES(seg_count, n * 4);
    MM_BYTES(ncand, n * 4);
    MM_CUDA(cudaMemcpyAsync(cx.cuts, cuts, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    int64_t h2d = n * 4;
    if (from_image) {
        MM_BYTES(starts, n * 4);
        MM_BYTES(lens, n * 4);
        MM_CUDA(cudaMemcpyAsync(cx.starts, starts, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        MM_CUDA(cudaMemcpyAsync(cx.lens, lens, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        h2d += n * 8;
    } else {
        MM_BYTES(win, std::max<int64_t>(total_len, 1));
        MM_BYTES(off, (n + 1) * 8);
        MM_CUDA(cudaMemcpyAsync(cx.win, windows, (size_t)total_len, cudaMemcpyHostToDevice, st));
        MM_CUDA(cudaMemcpyAsync(cx.off, offsets, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
        h2d += total_len + (n + 1) * 8;
    }
    const char *force_general = getenv("GUIDO_MMEJ_GENERAL");  // "1": skip the bit-parallel kernel (tests, comparisons)
    bool fast_pass = !(force_general && force_general[0] == '1');
    // letters of the windows the general kernel has to redo when the input is an image
    std::vector<char> redo_letters;
    std::vector<int64_t> redo_offsets;
    std::vector<int32_t> redo_cuts;
    std::vector<int> todo;  // empty = all windows
    uint32_t scratch_cap = 8192;
    int64_t out_cap = std::max<int64_t>(std::max<int64_t>(cap, 320 * n), 1 << 16);
    std::vector<uint32_t> result;           // kept tuples of all passes, completion order
    std::vector<long long> final_start((size_t)n, -1);
    bool done = false;
    for (int pass = 0; pass < 6 && !done; pass++) {
        const int n_items = todo.empty() ? (int)n : (int)todo.size();
        int grid = (int)std::min<int64_t>((n_items + kMmejWarps - 1) / kMmejWarps, (int64_t)ds.sm_count * 8);
        grid = std::max(grid, 1);
        const size_t per_warp = (size_t)scratch_cap + scratch_cap / 32;
        MM_BYTES(scratch, per_warp * 4 * (size_t)grid * kMmejWarps);
        MM_BYTES(out, out_cap * 4);
        MmejArgs a;
        a.windows = (const char *)cx.win; a.offsets = (const long long *)cx.off; a.cuts = (const int *)cx.cuts; a.n = (int)n;
        a.scratch = (uint32_t *)cx.scratch; a.cap = scratch_cap; a.out = (uint32_t *)cx.out; a.out_cap = (unsigned long long)out_cap;
        a.cursor = d_cursor; a.seg_start = (long long *)cx.seg_start; a.seg_count = (int *)cx.seg_count; a.n_cand = (int *)cx.ncand;
        a.todo = nullptr; a.n_todo = 0;
        const bool fast = fast_pass && todo.empty();
        if (!todo.empty()) {
            if (from_image) {
                // the general kernel reads letters: the batch of this pass becomes windows 0 .. n_todo-1
                a.n = (int)todo.size();
                MM_BYTES(win, std::max<size_t>(redo_letters.size(), 1));
                MM_BYTES(off, (todo.size() + 1) * 8);
                MM_CUDA(cudaMemcpyAsync(cx.win, redo_letters.data(), redo_letters.size(), cudaMemcpyHostToDevice, st));
                MM_CUDA(cudaMemcpyAsync(cx.off, redo_offsets.data(), (todo.size() + 1) * 8, cudaMemcpyHostToDevice, st));
                MM_CUDA(cudaMemcpyAsync(cx.cuts, redo_cuts.data(), todo.size() * 4, cudaMemcpyHostToDevice, st));
                a.windows = (const char *)cx.win; a.offsets = (const long long *)cx.off;
                h2d += (int64_t)redo_letters.size() + (int64_t)todo.size() * 12;
            } else {
                MM_BYTES(todo, todo.size() * 4);
                MM_CUDA(cudaMemcpyAsync(cx.todo, todo.data(), todo.size() * 4, cudaMemcpyHostToDevice, st));
                a.todo = (const int *)cx.todo; a.n_todo = (int)todo.size();
            }
        }
        MM_CUDA(cudaMemsetAsync(d_cursor, 0, 8, st));
        MM_CUDA(cudaEventRecord(ds.ev[1], st));
        if (fast) {
            FastArgs fa;
            fa.m = a;
            if (from_image) {
                const DeviceState *ids = img->dev;
                fa.m.windows = nullptr; fa.m.offsets = nullptr;
                fa.pv.planes = ids->d_planes; fa.pv.blk_lo = ids->blk_lo; fa.pv.blk_n = ids->blk_n;
                fa.pv.run_start = ids->d_run_start; fa.pv.run_end = ids->d_run_end; fa.pv.run_kind = ids->d_run_kind;
                fa.pv.tile_first_run = ids->d_tile_first_run; fa.pv.n_runs = ids->n_runs;
                fa.starts = (const uint32_t *)cx.starts; fa.lens = (const int *)cx.lens;
            } else {
                fa.pv = PlanesView{}; fa.starts = nullptr; fa.lens = nullptr;
            }
            mmej_fast_kernel<<<grid, kMmejWarps * 32, 0, st>>>(fa);
        } else {
            mmej_kernel<<<grid, kMmejWarps * 32, 0, st>>>(a);
        }
        MM_CUDA(cudaGetLastError());
        launches++;
        MM_CUDA(cudaEventRecord(ds.ev[2], st));
        const int n_seen = from_image && !todo.empty() ? (int)todo.size() : (int)n;  // windows the device arrays describe
        unsigned long long cursor = 0;
        MM_CUDA(cudaMemcpyAsync(&cursor, d_cursor, 8, cudaMemcpyDeviceToHost, st));
        MM_CUDA(cudaMemcpyAsync(seg_start.data(), cx.seg_start, (size_t)n_seen * 8, cudaMemcpyDeviceToHost, st));
        MM_CUDA(cudaMemcpyAsync(seg_count.data(), cx.seg_count, (size_t)n_seen * 4, cudaMemcpyDeviceToHost, st));
        MM_CUDA(cudaMemcpyAsync(ncand.data(), cx.ncand, (size_t)n_seen * 4, cudaMemcpyDeviceToHost, st));
        MM_CUDA(cudaStreamSynchronize(st));
        float ms = 0;
        cudaEventElapsedTime(&ms, ds.ev[1], ds.ev[2]);
        kernel_ms += ms;
        if ((int64_t)cursor > out_cap) {  // output buffer too small: same pass again, exact size
            out_cap = (int64_t)cursor;
            pass--;
            if (launches > 10) { set_error("internal: mmej output sizing did not converge"); return GUIDO_ERR_CUDA; }
            continue;
        }
        if (pass == 0 && todo.empty() && cursor <= 0xffffffffull) {
            // the usual case, every window done by the first launch: segments are brought into window
            // order on the device (scan of the counts, one warp per window) and go to the caller's
            // buffer in ONE copy -- no staging vector, no per-window memcpy on the host
            bool all_done = true;
            for (int64_t g = 0; g < n && all_done; g++) all_done = seg_start[(size_t)g] >= 0;
            if (all_done) {
                const uint32_t n1 = (uint32_t)n + 1;
                MM_BYTES(offs, ((int64_t)n1 + (n1 + 4095) / 4096 + 8) * 4);
                MM_BYTES(fin, std::max<int64_t>((int64_t)cursor, 1) * 4);
                uint32_t *d_offs = (uint32_t *)cx.offs, *d_sums = d_offs + n1;
                mmej_counts_kernel<<<(n1 + 255) / 256, 256, 0, st>>>((const int *)cx.seg_count, (int)n, d_offs);
                if ((rc = exclusive_scan_u32(d_offs, n1, d_sums, nullptr, st))) return rc;
                mmej_order_kernel<<<std::min<int64_t>((n + 7) / 8, (int64_t)ds.sm_count * 16), 256, 0, st>>>(
                    (const uint32_t *)cx.out, (const long long *)cx.seg_start, d_offs, (int)n, (uint32_t *)cx.fin, cursor);
                MM_CUDA(cudaGetLastError());
                launches += 5;
                std::vector<uint32_t> offs32((size_t)n1);
                MM_CUDA(cudaMemcpyAsync(offs32.data(), d_offs, (size_t)n1 * 4, cudaMemcpyDeviceToHost, st));
                if ((int64_t)cursor <= cap && cursor)
                    MM_CUDA(cudaMemcpyAsync(out_tuples, cx.fin, (size_t)cursor * 4, cudaMemcpyDeviceToHost, st));
                MM_CUDA(cudaEventRecord(ds.ev[3], st));
                MM_CUDA(cudaStreamSynchronize(st));
                for (int64_t g = 0; g <= n; g++) out_offsets[g] = offs32[(size_t)g];
                if (n_cand) memcpy(n_cand, ncand.data(), (size_t)n * 4);
                if (stats) {
                    float t = 0;
                    cudaEventElapsedTime(&t, ds.ev[0], ds.ev[3]);
                    stats->kernel_ms = kernel_ms; stats->total_ms = t; stats->n_launches = launches;
                    stats->n_sites = n; stats->n_pairs = 0;
                    stats->bytes_h2d = h2d + 8; stats->bytes_d2h = (int64_t)cursor * 4 + n * 20;
                }
                if ((int64_t)cursor > cap) { set_error("output capacity too small: need %lld tuples", (long long)cursor); return GUIDO_ERR_CAPACITY; }
                return GUIDO_OK;
            }
        }
        const size_t base = result.size();
        result.resize(base + (size_t)cursor);
        if (cursor) MM_CUDA(cudaMemcpyAsync(result.data() + base, cx.out, (size_t)cursor * 4, cudaMemcpyDeviceToHost, st));
        MM_CUDA(cudaStreamSynchronize(st));
        // slot = index in the device arrays of this pass, g = window of the call
        std::vector<int> next;
        bool grow = false;
        auto handle = [&](int slot, int g) {
            if (seg_start[(size_t)slot] < 0) { next.push_back(g); grow |= seg_start[(size_t)slot] == -1; return; }
            final_start[(size_t)g] = (long long)base + seg_start[(size_t)slot];
            final_count[(size_t)g] = seg_count[(size_t)slot];
            final_ncand[(size_t)g] = ncand[(size_t)slot];
        };
        if (todo.empty()) for (int g = 0; g < (int)n; g++) handle(g, g);
        else for (size_t t = 0; t < todo.size(); t++) handle(from_image ? (int)t : todo[t], todo[t]);
        todo.swap(next);
        if (todo.empty()) { done = true; break; }
        if (from_image) {
            // letters of the windows still open, from the host copy of the image
            redo_letters.clear(); redo_offsets.assign(1, 0); redo_cuts.clear();
            for (int g : todo) {
                const size_t at = redo_letters.size();
                redo_letters.resize(at + (size_t)lens[g]);
                std::vector<int64_t> gp((size_t)lens[g]);
                for (int i = 0; i < lens[g]; i++) gp[(size_t)i] = (int64_t)starts[g] + i;
                if ((rc = guido_image_fetch_windows(img, gp.data(), lens[g], 1, redo_letters.data() + at))) return rc;
                redo_offsets.push_back((int64_t)redo_letters.size());
                redo_cuts.push_back(cuts[g]);
            }
        }
        if (!fast && grow) scratch_cap *= 16;  // 8192 -> 131072 -> 2M candidates per warp for pathological repeats
        out_cap = std::max<int64_t>(out_cap, (int64_t)scratch_cap * (int64_t)todo.size());
    }
    int64_t total_kept = 0;
    if (done) {
        for (int64_t g = 0; g < n; g++) { out_offsets[g] = total_kept; total_kept += final_count[(size_t)g]; }
        out_offsets[n] = total_kept;
        if (total_kept <= cap)
            for (int64_t g = 0; g < n; g++)
                if (final_count[(size_t)g])
                    memcpy(out_tuples + out_offsets[g], result.data() + final_start[(size_t)g], (size_t)final_count[(size_t)g] * 4);
        if (n_cand) memcpy(n_cand, final_ncand.data(), (size_t)n * 4);
    }
    MM_CUDA(cudaEventRecord(ds.ev[3], st));
    MM_CUDA(cudaStreamSynchronize(st));
    if (stats) {
        float t = 0;
        cudaEventElapsedTime(&t, ds.ev[0], ds.ev[3]);
        stats->kernel_ms = kernel_ms; stats->total_ms = t; stats->n_launches = launches;
        stats->n_sites = n; stats->n_pairs = 0;
        stats->bytes_h2d = h2d + 8; stats->bytes_d2h = total_kept * 4 + n * 16;
    }
    if (!done) { set_error("mmej: %zu windows exceed the candidate scratch", todo.size()); return GUIDO_ERR_NOMEM; }
    if (total_kept > cap) { set_error("output capacity too small: need %lld tuples", (long long)total_kept); return GUIDO_ERR_CAPACITY; }
    return GUIDO_OK;
}
#undef MM_CUDA
#undef MM_BYTES

}  // namespace

extern "C" int guido_mmej_batch(const char *windows, const int64_t *offsets, const int32_t *cuts, int64_t n,
                                int32_t device, uint32_t *out_tuples, int64_t cap, int64_t *out_offsets,
                                int32_t *n_cand, guido_stats *stats) {
    GUIDO_API_RANGE();
    if (n < 0 || (n > 0 && (!windows || !offsets || !cuts)) || !out_offsets) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    if (n > 0x7fffffff) { set_error("too many windows"); return GUIDO_ERR_ARG; }
    for (int64_t g = 0; g < n; g++) {
        int64_t len = offsets[g + 1] - offsets[g];
        int64_t cut = std::min<int64_t>(std::max<int64_t>(cuts[g], 0), len);
        if (len < 0 || cut > kMaxSide || len - cut > kMaxSide) {
            set_error("window %lld: each side of the cut may hold at most %d letters", (long long)g, kMaxSide);
            return GUIDO_ERR_ARG;
        }
    }
    out_offsets[0] = 0;
    if (n == 0) return GUIDO_OK;
    std::lock_guard<std::mutex> lock(g_mmej_mu);
    MmejCtx *cx = nullptr;
    int rc = mmej_ctx(device, &cx);
    if (rc) return rc;
    return mmej_run(*cx, windows, offsets, nullptr, nullptr, nullptr, cuts, n, out_tuples, cap, out_offsets, n_cand, stats);
}

extern "C" int guido_mmej_sites(const guido_image *img, const uint32_t *starts, const int32_t *lens, const int32_t *cuts, int64_t n,
                                uint32_t *out_tuples, int64_t cap, int64_t *out_offsets, int32_t *n_cand, guido_stats *stats) {
    GUIDO_API_RANGE();
    if (!img || n < 0 || (n > 0 && (!starts || !lens || !cuts)) || !out_offsets) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    if (!img->dev) { set_error("image is not resident on a device"); return GUIDO_ERR_ARG; }
    if (n > 0x7fffffff) { set_error("too many windows"); return GUIDO_ERR_ARG; }
    const DeviceState *ids = img->dev;
    for (int64_t g = 0; g < n; g++) {
        const int64_t len = lens[g], cut = std::min<int64_t>(std::max<int64_t>(cuts[g], 0), len);
        if (len < 0 || cut > kMaxSide || len - cut > kMaxSide) {
            set_error("window %lld: each side of the cut may hold at most %d letters", (long long)g, kMaxSide);
            return GUIDO_ERR_ARG;
        }
        const int64_t b0 = (int64_t)starts[g] >> 6, b1 = ((int64_t)starts[g] + len + 63) >> 6;
        if (b0 < ids->blk_lo || b1 > ids->blk_lo + ids->blk_n) {
            set_error("window %lld lies outside the resident part of the image", (long long)g);
            return GUIDO_ERR_ARG;
        }
    }
    out_offsets[0] = 0;
    if (n == 0) return GUIDO_OK;
    std::lock_guard<std::mutex> lock(g_mmej_mu);
    MmejCtx *cx = nullptr;
    int rc = mmej_ctx(ids->device, &cx);
    if (rc) return rc;
    return mmej_run(*cx, nullptr, nullptr, img, starts, lens, cuts, n, out_tuples, cap, out_offsets, n_cand, stats);
}
