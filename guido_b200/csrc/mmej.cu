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

}  // namespace guido

using namespace guido;

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
    // Stream, events and device buffers are kept per device between calls (grown on demand): a
    // locus-sized batch runs for milliseconds, a context and nine allocations per call cost as much
    struct Ctx {
        DeviceState ds;
        void *win = nullptr, *off = nullptr, *cuts = nullptr, *seg_start = nullptr, *seg_count = nullptr, *ncand = nullptr,
             *todo = nullptr, *scratch = nullptr, *out = nullptr;
        int64_t win_cap = 0, off_cap = 0, cuts_cap = 0, seg_start_cap = 0, seg_count_cap = 0, ncand_cap = 0, todo_cap = 0,
                scratch_cap = 0, out_cap = 0;
    };
    static std::mutex mu;
    static Ctx *cached[64] = {};
    std::lock_guard<std::mutex> lock(mu);
    if (device < 0 || device >= 64) { set_error("device %d out of range", (int)device); return GUIDO_ERR_ARG; }
    if (!cached[device]) {
        Ctx *c = new Ctx();
        int rc0 = device_init(&c->ds, device);
        if (rc0) { delete c; return rc0; }
        cached[device] = c;
    }
    Ctx &cx = *cached[device];
    DeviceState &ds = cx.ds;
    int rc;
    GUIDO_CUDA(cudaSetDevice(device));
    const int64_t total_len = offsets[n];
    unsigned long long *d_cursor = ds.d_counts;
    std::vector<long long> seg_start((size_t)n);
    std::vector<int> seg_count((size_t)n), ncand((size_t)n);
    std::vector<uint32_t> h_out;
    int launches = 0;
    float kernel_ms = 0;
    auto cleanup = [&]() {};  // the buffers stay with the cached context
#define MM_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_error("CUDA error %s at %s:%d", cudaGetErrorName(e_), __FILE__, __LINE__); cleanup(); return GUIDO_ERR_CUDA; } } while (0)
#define MM_BYTES(p, need) do { if ((rc = ensure_bytes(&cx.p, &cx.p##_cap, (int64_t)(need)))) return rc; } while (0)
    cudaStream_t st = ds.stream;
    MM_CUDA(cudaEventRecord(ds.ev[0], st));
    MM_BYTES(win, std::max<int64_t>(total_len, 1));
    MM_BYTES(off, (n + 1) * 8);
    MM_BYTES(cuts, n * 4);
    MM_BYTES(seg_start, n * 8);
    MM_BYTES(seg_count, n * 4);
    MM_BYTES(ncand, n * 4);
    char *d_win = (char *)cx.win; long long *d_off = (long long *)cx.off, *d_seg_start = (long long *)cx.seg_start;
    int *d_cuts = (int *)cx.cuts, *d_seg_count = (int *)cx.seg_count, *d_ncand = (int *)cx.ncand, *d_todo = nullptr;
    uint32_t *d_scratch = nullptr, *d_out = nullptr;
    MM_CUDA(cudaMemcpyAsync(d_win, windows, (size_t)total_len, cudaMemcpyHostToDevice, st));
    MM_CUDA(cudaMemcpyAsync(d_off, offsets, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
    MM_CUDA(cudaMemcpyAsync(d_cuts, cuts, (size_t)n * 4, cudaMemcpyHostToDevice, st));

    std::vector<int> todo;  // empty = all windows
    uint32_t scratch_cap = 8192;
    int64_t out_cap = std::max<int64_t>(std::max<int64_t>(cap, 320 * n), 1 << 16);
    int64_t total_kept = 0;
    std::vector<uint32_t> result;           // kept tuples of all passes, completion order
    std::vector<long long> final_start((size_t)n, -1);
    for (int pass = 0; pass < 4; pass++) {
        const int n_items = todo.empty() ? (int)n : (int)todo.size();
        int grid = std::min<int64_t>((n_items + kMmejWarps - 1) / kMmejWarps, (int64_t)ds.sm_count * 8);
        grid = std::max(grid, 1);
        const size_t per_warp = (size_t)scratch_cap + scratch_cap / 32;
        MM_BYTES(scratch, per_warp * 4 * (size_t)grid * kMmejWarps);
        MM_BYTES(out, out_cap * 4);
        d_scratch = (uint32_t *)cx.scratch; d_out = (uint32_t *)cx.out;
        if (!todo.empty()) {
            MM_BYTES(todo, todo.size() * 4);
            d_todo = (int *)cx.todo;
            MM_CUDA(cudaMemcpyAsync(d_todo, todo.data(), todo.size() * 4, cudaMemcpyHostToDevice, st));
        }
        MM_CUDA(cudaMemsetAsync(d_cursor, 0, 8, st));
        MmejArgs a;
        a.windows = d_win; a.offsets = d_off; a.cuts = d_cuts; a.n = (int)n;
        a.scratch = d_scratch; a.cap = scratch_cap; a.out = d_out; a.out_cap = (unsigned long long)out_cap;
        a.cursor = d_cursor; a.seg_start = d_seg_start; a.seg_count = d_seg_count; a.n_cand = d_ncand;
        a.todo = todo.empty() ? nullptr : d_todo; a.n_todo = (int)todo.size();
        MM_CUDA(cudaEventRecord(ds.ev[1], st));
        mmej_kernel<<<grid, kMmejWarps * 32, 0, st>>>(a);
        MM_CUDA(cudaGetLastError());
        launches++;
        MM_CUDA(cudaEventRecord(ds.ev[2], st));
        unsigned long long cursor = 0;
        MM_CUDA(cudaMemcpyAsync(&cursor, d_cursor, 8, cudaMemcpyDeviceToHost, st));
        MM_CUDA(cudaMemcpyAsync(seg_start.data(), d_seg_start, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        MM_CUDA(cudaMemcpyAsync(seg_count.data(), d_seg_count, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        MM_CUDA(cudaMemcpyAsync(ncand.data(), d_ncand, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        MM_CUDA(cudaStreamSynchronize(st));
        float ms = 0;
        cudaEventElapsedTime(&ms, ds.ev[1], ds.ev[2]);
        kernel_ms += ms;
        if ((int64_t)cursor > out_cap) {  // output buffer too small: same pass again, exact size
            out_cap = (int64_t)cursor;
            pass--;
            if (launches > 8) { set_error("internal: mmej output sizing did not converge"); cleanup(); return GUIDO_ERR_CUDA; }
            continue;
        }
        size_t base = result.size();
        result.resize(base + (size_t)cursor);
        if (cursor) MM_CUDA(cudaMemcpyAsync(result.data() + base, d_out, (size_t)cursor * 4, cudaMemcpyDeviceToHost, st));
        MM_CUDA(cudaStreamSynchronize(st));
        std::vector<int> next;
        auto handle = [&](int g) {
            if (seg_start[(size_t)g] < 0) next.push_back(g);
            else final_start[(size_t)g] = (long long)base + seg_start[(size_t)g];
        };
        if (todo.empty()) for (int g = 0; g < (int)n; g++) handle(g);
        else for (int g : todo) handle(g);
        // seg_count of finished windows must survive later passes
        static thread_local std::vector<int> final_count;
        if (pass == 0) final_count.assign((size_t)n, 0);
        if (todo.empty()) { for (int g = 0; g < (int)n; g++) if (seg_start[(size_t)g] >= 0) final_count[(size_t)g] = seg_count[(size_t)g]; }
        else for (int g : todo) if (seg_start[(size_t)g] >= 0) final_count[(size_t)g] = seg_count[(size_t)g];
        if (n_cand) memcpy(n_cand, ncand.data(), (size_t)n * 4);
        todo.swap(next);
        if (todo.empty()) {
            total_kept = 0;
            for (int64_t g = 0; g < n; g++) { out_offsets[g] = total_kept; total_kept += final_count[(size_t)g]; }
            out_offsets[n] = total_kept;
            if (total_kept <= cap)
                for (int64_t g = 0; g < n; g++)
                    if (final_count[(size_t)g])
                        memcpy(out_tuples + out_offsets[g], result.data() + final_start[(size_t)g], (size_t)final_count[(size_t)g] * 4);
            break;
        }
        scratch_cap *= 16;  // 8192 -> 131072 -> 2M candidates per warp for pathological repeats
        out_cap = std::max<int64_t>(out_cap, (int64_t)scratch_cap * (int64_t)todo.size());
    }
    MM_CUDA(cudaEventRecord(ds.ev[3], st));
    MM_CUDA(cudaStreamSynchronize(st));
    if (stats) {
        float t = 0;
        cudaEventElapsedTime(&t, ds.ev[0], ds.ev[3]);
        stats->kernel_ms = kernel_ms; stats->total_ms = t; stats->n_launches = launches;
        stats->n_sites = n; stats->n_pairs = 0;
        stats->bytes_h2d = total_len + n * 12 + 8; stats->bytes_d2h = total_kept * 4 + n * 16;
    }
    cleanup();
#undef MM_CUDA
#undef MM_BYTES
    if (!todo.empty()) { set_error("mmej: %zu windows exceed the candidate scratch", todo.size()); return GUIDO_ERR_NOMEM; }
    if (total_kept > cap) { set_error("output capacity too small: need %lld tuples", (long long)total_kept); return GUIDO_ERR_CAPACITY; }
    return GUIDO_OK;
}
