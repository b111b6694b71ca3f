// sort.cu -- canonical order of the off-target hit list: (guide, gpos, strand), on the device.
//
// The search kernels emit hits in completion order (atomic cursor).  The host-buffer C-ABI call
// promises a deterministic order, so 1/2/4/8-GPU runs and repeated runs are byte-identical.  This
// is a plain key-value radix sort (CUB DeviceRadixSort, a library call -- not a hot-path kernel):
// key = guide << 33 | gpos << 1 | strand, value = {hi, lo}.
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

// sort d_hits[0..n) in place; scratch comes from ds->d_out[1]
int sort_hits_device(DeviceState *ds, guido_ot_hit *d_hits, uint64_t n, int64_t n_guides, cudaStream_t stream,
                     int *n_launches) {
    if (n < 2) return GUIDO_OK;
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
