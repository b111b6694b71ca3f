// capi.cu -- the extern "C" entry points that touch the device (see include/guido_b200.h).
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "device.cuh"

namespace guido {

int ensure_bytes(void **p, int64_t *cap, int64_t need) {
    if (need <= *cap) return GUIDO_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    int64_t want = std::max<int64_t>(need, 1 << 16);
    GUIDO_CUDA(cudaMalloc(p, (size_t)want));
    *cap = want;
    return GUIDO_OK;
}

int ensure_pinned(DeviceState *ds, int64_t need) {
    if (need <= ds->pinned_cap) return GUIDO_OK;
    if (ds->h_pinned) cudaFreeHost(ds->h_pinned);
    ds->h_pinned = nullptr; ds->pinned_cap = 0;
    int64_t want = std::max<int64_t>(need, 1 << 20);
    GUIDO_CUDA(cudaMallocHost(&ds->h_pinned, (size_t)want));
    ds->pinned_cap = want;
    return GUIDO_OK;
}

int device_init(DeviceState *ds, int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        set_error("no CUDA device: guido_b200 has no CPU fallback");
        return GUIDO_ERR_CUDA;
    }
    if (device < 0 || device >= n) { set_error("device %d out of range (0..%d)", device, n - 1); return GUIDO_ERR_ARG; }
    GUIDO_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    GUIDO_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return GUIDO_ERR_CUDA;
    }
    ds->device = device;
    ds->sm_count = prop.multiProcessorCount;
    GUIDO_CUDA(cudaStreamCreateWithFlags(&ds->stream, cudaStreamNonBlocking));
    for (auto &e : ds->ev) GUIDO_CUDA(cudaEventCreate(&e));
    GUIDO_CUDA(cudaMalloc(&ds->d_counts, 8 * sizeof(unsigned long long)));
    GUIDO_CUDA(cudaMemset(ds->d_counts, 0, 8 * sizeof(unsigned long long)));  // [7]: look-back gave up (pam_scan.cu)
    GUIDO_CUDA(cudaMalloc(&ds->d_tile_counter, 4 * sizeof(unsigned int)));
    return GUIDO_OK;
}

void kev_begin(DeviceState *ds, cudaStream_t stream) {
    cudaEvent_t *p = ds->kev[ds->kev_count % DeviceState::kKev];
    if (!p[0]) { cudaEventCreate(&p[0]); cudaEventCreate(&p[1]); }  // created on first use
    cudaEventRecord(p[0], stream);
}

void kev_end(DeviceState *ds, cudaStream_t stream) {
    cudaEventRecord(ds->kev[ds->kev_count % DeviceState::kKev][1], stream);
    ds->kev_count++;
}

void device_release(guido_image *img) {
    DeviceState *ds = img->dev;
    if (!ds) return;
    if (ds->device >= 0) {
        cudaSetDevice(ds->device);
        if (ds->stream) cudaStreamSynchronize(ds->stream);
        cudaFree(ds->d_planes); cudaFree(ds->d_run_start); cudaFree(ds->d_run_end); cudaFree(ds->d_run_kind); cudaFree(ds->d_tile_first_run);
        cudaFree(ds->d_desc); cudaFree(ds->d_counts); cudaFree(ds->d_tile_counter);
        cudaFree(ds->d_out[0]); cudaFree(ds->d_out[1]); cudaFree(ds->d_guides); cudaFree(ds->d_flags);
        cudaFree(ds->d_index); cudaFree(ds->d_bucket);
        release_site_table(ds);
        cudaFree(ds->d_site_off); cudaFree(ds->d_tickets);
        if (ds->h_pinned) cudaFreeHost(ds->h_pinned);
        for (auto &e : ds->ev) if (e) cudaEventDestroy(e);
        for (auto &p : ds->kev) { if (p[0]) cudaEventDestroy(p[0]); if (p[1]) cudaEventDestroy(p[1]); }
        for (auto &e : ds->ev_slice) if (e) cudaEventDestroy(e);
        if (ds->build_stream) cudaStreamDestroy(ds->build_stream);
        if (ds->join_stream) cudaStreamDestroy(ds->join_stream);
        if (ds->stream) cudaStreamDestroy(ds->stream);
    }
    delete ds;
    img->dev = nullptr;
}

static PlanesView view_of(const DeviceState *ds) {
    PlanesView pv;
    pv.planes = ds->d_planes; pv.blk_lo = ds->blk_lo; pv.blk_n = ds->blk_n;
    pv.run_start = ds->d_run_start; pv.run_end = ds->d_run_end; pv.run_kind = ds->d_run_kind; pv.tile_first_run = ds->d_tile_first_run;
    pv.n_runs = ds->n_runs;
    return pv;
}

// (re)loads the image's shard into `ds` (initialised for a device); buffers are reused when they
// are large enough, so repeated uploads through one DeviceState cost copies, not allocations
static int load_shard(guido_image *img, DeviceState *ds, int shard_ix, int n_shards) {
    GUIDO_CUDA(cudaSetDevice(ds->device));
    release_site_table(ds);  // whatever was tabled belongs to the previous content
    // contiguous shards of equal size, aligned to scan tiles (1024 blocks = 65536 bases); a shard
    // owns the PAM starts inside it and keeps a halo of 4 blocks (>= 20 + P bases) on either side
    guido_shard_bounds(img->padded_bases, shard_ix, n_shards, &ds->shard_lo, &ds->shard_hi);
    const int64_t b0 = ds->shard_lo / 64, b1 = ds->shard_hi / 64;
    ds->blk_lo = std::max<int64_t>((b0 & ~(int64_t)3) - 4, 0);   // multiple of 4: 256-bit loads stay aligned
    int64_t blk_hi = std::min<int64_t>(b1 + 4, img->n_blocks);
    ds->blk_n = blk_hi - ds->blk_lo;
    int rc;
    if ((rc = ensure_bytes((void **)&ds->d_planes, &ds->planes_cap, (ds->blk_n + 1) * (int64_t)sizeof(ulonglong2)))) return rc;
    GUIDO_CUDA(cudaMemsetAsync(ds->d_planes + ds->blk_n, 0, sizeof(ulonglong2), ds->stream));
    GUIDO_CUDA(cudaMemcpyAsync(ds->d_planes, img->planes.data() + 2 * ds->blk_lo,
                               (size_t)ds->blk_n * sizeof(ulonglong2), cudaMemcpyHostToDevice, ds->stream));
    ds->n_runs = (int)img->runs.size();
    std::vector<uint32_t> rs(ds->n_runs), re(ds->n_runs);
    std::vector<uint8_t> rk(ds->n_runs);
    for (int i = 0; i < ds->n_runs; i++) { rs[i] = img->runs[i].start; re[i] = img->runs[i].end; rk[i] = img->runs[i].kind; }
    const int64_t nr = std::max(ds->n_runs, 1);
    if ((rc = ensure_bytes((void **)&ds->d_run_start, &ds->run_start_cap, nr * 4))) return rc;
    if ((rc = ensure_bytes((void **)&ds->d_run_end, &ds->run_end_cap, nr * 4))) return rc;
    if ((rc = ensure_bytes((void **)&ds->d_run_kind, &ds->run_kind_cap, nr))) return rc;
    GUIDO_CUDA(cudaMemcpyAsync(ds->d_run_start, rs.data(), (size_t)ds->n_runs * 4, cudaMemcpyHostToDevice, ds->stream));
    GUIDO_CUDA(cudaMemcpyAsync(ds->d_run_end, re.data(), (size_t)ds->n_runs * 4, cudaMemcpyHostToDevice, ds->stream));
    GUIDO_CUDA(cudaMemcpyAsync(ds->d_run_kind, rk.data(), (size_t)ds->n_runs, cudaMemcpyHostToDevice, ds->stream));
    // first run whose end lies beyond (tile start - 32), one entry per 65536-base tile
    const int64_t n_rt = img->padded_bases / 65536 + 2;
    std::vector<uint32_t> first((size_t)n_rt);
    size_t ri = 0;
    for (int64_t t = 0; t < n_rt; t++) {
        const int64_t from = std::max<int64_t>(t * 65536 - 32, 0);
        while (ri < img->runs.size() && (int64_t)img->runs[ri].end <= from) ri++;
        // bit 31: some run touches [tile start - 32, tile end + 32) -- the kernels' slow-path flag
        const bool touched = ri < img->runs.size() && (int64_t)img->runs[ri].start < (t + 1) * 65536 + 32;
        first[(size_t)t] = (uint32_t)ri | (touched ? 0x80000000u : 0u);
    }
    if ((rc = ensure_bytes((void **)&ds->d_tile_first_run, &ds->tile_first_run_cap, n_rt * 4))) return rc;
    GUIDO_CUDA(cudaMemcpyAsync(ds->d_tile_first_run, first.data(), (size_t)n_rt * 4, cudaMemcpyHostToDevice, ds->stream));
    GUIDO_CUDA(cudaStreamSynchronize(ds->stream));  // the host vectors above go out of scope
    return GUIDO_OK;
}

static int upload_impl(guido_image *img, int device, int shard_ix, int n_shards) {
    if (!img || n_shards < 1 || shard_ix < 0 || shard_ix >= n_shards) { set_error("bad shard arguments"); return GUIDO_ERR_ARG; }
    if (!img->dev || img->dev->device != device) {  // a handle that moves to another device starts over
        device_release(img);
        auto *ds = new DeviceState();
        img->dev = ds;
        int rc = device_init(ds, device);
        if (rc) { device_release(img); return rc; }
    }
    int rc = load_shard(img, img->dev, shard_ix, n_shards);
    if (rc) device_release(img);
    return rc;
}

static int need_device(guido_image *img) {
    if (!img || !img->dev || img->dev->device < 0) {
        set_error("image is not uploaded to a device (call guido_image_upload); there is no CPU fallback");
        return GUIDO_ERR_CUDA;
    }
    GUIDO_CUDA(cudaSetDevice(img->dev->device));
    return GUIDO_OK;
}

// run one scan, copy counts (and as many positions as fit) back to the host
static int scan_to_host(guido_image *img, const PamSpec &pam, int mode, const ScanRange &r,
                        uint32_t *fw, int64_t cap_fw, int64_t *n_fw, uint32_t *rv, int64_t cap_rv,
                        int64_t *n_rv, guido_stats *stats) {
    DeviceState *ds = img->dev;
    int rc;
    int64_t span = (int64_t)r.hi - (int64_t)r.lo;
    int64_t dcap_f = std::min<int64_t>(std::max<int64_t>(cap_fw, 0), span), dcap_r = std::min<int64_t>(std::max<int64_t>(cap_rv, 0), span);
    if ((rc = ensure_bytes(&ds->d_out[0], &ds->out_cap[0], dcap_f * 4))) return rc;
    if ((rc = ensure_bytes(&ds->d_out[1], &ds->out_cap[1], dcap_r * 4))) return rc;
    int launches = 0;
    GUIDO_CUDA(cudaEventRecord(ds->ev[0], ds->stream));
    rc = launch_pam_scan(ds, view_of(ds), pam, mode, r, (uint32_t *)ds->d_out[0], (uint64_t)dcap_f,
                         (uint32_t *)ds->d_out[1], (uint64_t)dcap_r, ds->d_counts, ds->stream, &launches);
    if (rc) return rc;
    GUIDO_CUDA(cudaEventRecord(ds->ev[1], ds->stream));
    unsigned long long counts[2], gave_up = 0;
    GUIDO_CUDA(cudaMemcpyAsync(counts, ds->d_counts, sizeof counts, cudaMemcpyDeviceToHost, ds->stream));
    GUIDO_CUDA(cudaMemcpyAsync(&gave_up, ds->d_counts + 7, sizeof gave_up, cudaMemcpyDeviceToHost, ds->stream));
    GUIDO_CUDA(cudaStreamSynchronize(ds->stream));
    if (gave_up) {
        set_error("PAM scan: a tile never published its counts (look-back gave up)");
        return GUIDO_ERR_CUDA;
    }
    *n_fw = (int64_t)counts[0];
    *n_rv = (int64_t)counts[1];
    int64_t cf = std::min<int64_t>(*n_fw, dcap_f), cr = std::min<int64_t>(*n_rv, dcap_r);
    if (cf) GUIDO_CUDA(cudaMemcpyAsync(fw, ds->d_out[0], (size_t)cf * 4, cudaMemcpyDeviceToHost, ds->stream));
    if (cr) GUIDO_CUDA(cudaMemcpyAsync(rv, ds->d_out[1], (size_t)cr * 4, cudaMemcpyDeviceToHost, ds->stream));
    GUIDO_CUDA(cudaEventRecord(ds->ev[2], ds->stream));
    GUIDO_CUDA(cudaStreamSynchronize(ds->stream));
    if (stats) {
        float k = 0, t = 0;
        cudaEventElapsedTime(&k, ds->ev[0], ds->ev[1]);
        cudaEventElapsedTime(&t, ds->ev[0], ds->ev[2]);
        stats->kernel_ms = k; stats->total_ms = t; stats->n_launches = launches;
        stats->n_sites = *n_fw + *n_rv; stats->n_pairs = 0;
        stats->bytes_h2d = 0; stats->bytes_d2h = 16 + (cf + cr) * 4;
    }
    if (*n_fw > cap_fw || *n_rv > cap_rv) {
        set_error("output capacity too small: need %lld fw / %lld rv", (long long)*n_fw, (long long)*n_rv);
        return GUIDO_ERR_CAPACITY;
    }
    return GUIDO_OK;
}

}  // namespace guido

using namespace guido;

extern "C" {

int guido_device_count(void) {
    GUIDO_API_RANGE();
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    int ok = 0;
    for (int i = 0; i < n; i++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major >= 10) ok++;
    }
    return ok;
}

int guido_shard_bounds(int64_t padded_bases, int32_t shard_ix, int32_t n_shards, int64_t *lo, int64_t *hi) {
    GUIDO_API_RANGE();
    if (n_shards < 1 || shard_ix < 0 || shard_ix >= n_shards || !lo || !hi) { set_error("bad shard arguments"); return GUIDO_ERR_ARG; }
    const int64_t n_blocks = (padded_bases + 63) / 64;
    int64_t per = (n_blocks + n_shards - 1) / n_shards;
    per = (per + 1023) / 1024 * 1024;
    const int64_t b0 = std::min<int64_t>(per * shard_ix, n_blocks), b1 = std::min<int64_t>(b0 + per, n_blocks);
    *lo = b0 * 64;
    *hi = b1 * 64;
    return GUIDO_OK;
}

int guido_host_alloc(int64_t bytes, void **out) {
    GUIDO_API_RANGE();
    if (bytes < 0 || !out) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    *out = nullptr;
    GUIDO_CUDA(cudaHostAlloc(out, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocPortable));
    return GUIDO_OK;
}

int guido_host_free(void *p) {
    GUIDO_API_RANGE();
    if (p) GUIDO_CUDA(cudaFreeHost(p));
    return GUIDO_OK;
}

// ---- peer buffers: device memory another process on the node can write into (CUDA IPC) -----------
int guido_peer_alloc(int32_t device, int64_t bytes, void **d_ptr, void *handle64) {
    GUIDO_API_RANGE();
    if (bytes < 0 || !d_ptr || !handle64) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the C-ABI promises a 64-byte handle");
    GUIDO_CUDA(cudaSetDevice(device));
    *d_ptr = nullptr;
    GUIDO_CUDA(cudaMalloc(d_ptr, (size_t)std::max<int64_t>(bytes, 256)));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *d_ptr);
    if (e != cudaSuccess) {
        cudaFree(*d_ptr); *d_ptr = nullptr;
        set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
        return GUIDO_ERR_CUDA;
    }
    memcpy(handle64, &h, 64);
    return GUIDO_OK;
}

int guido_peer_open(int32_t device, const void *handle64, void **d_ptr) {
    GUIDO_API_RANGE();
    if (!handle64 || !d_ptr) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    GUIDO_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    GUIDO_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GUIDO_OK;
}

int guido_peer_close(void *d_ptr) {
    GUIDO_API_RANGE();
    if (d_ptr) GUIDO_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return GUIDO_OK;
}

int guido_peer_free(void *d_ptr) {
    GUIDO_API_RANGE();
    if (d_ptr) GUIDO_CUDA(cudaFree(d_ptr));
    return GUIDO_OK;
}

int guido_copy_dev(void *dst, const void *src, int64_t bytes, void *stream) {
    GUIDO_API_RANGE();
    if (bytes < 0 || (bytes && (!dst || !src))) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    if (bytes) GUIDO_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return GUIDO_OK;
}

int guido_peer_send(void *const *dst_ptrs, void *const *streams, int32_t n, int32_t first, int64_t at_bytes, const void *src,
                    int64_t bytes, void *after_event) {
    GUIDO_API_RANGE();
    if (n < 0 || (n && (!dst_ptrs || !streams)) || bytes < 0 || at_bytes < 0) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    for (int32_t k = 0; k < n && bytes; k++) {
        const int32_t r = (first + k) % n;  // start with `first` (the own buffer), then round the ring
        cudaStream_t st = (cudaStream_t)streams[r];
        if (after_event) GUIDO_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)after_event, 0));
        GUIDO_CUDA(cudaMemcpyAsync((char *)dst_ptrs[r] + at_bytes, src, (size_t)bytes, cudaMemcpyDefault, st));
    }
    return GUIDO_OK;
}

int guido_streams_join(void *const *streams, void *const *events, int32_t n, void *onto_stream) {
    GUIDO_API_RANGE();
    if (n < 0 || (n && (!streams || !events))) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    for (int32_t k = 0; k < n; k++) {
        GUIDO_CUDA(cudaEventRecord((cudaEvent_t)events[k], (cudaStream_t)streams[k]));
        GUIDO_CUDA(cudaStreamWaitEvent((cudaStream_t)onto_stream, (cudaEvent_t)events[k], 0));
    }
    return GUIDO_OK;
}

int guido_kernel_times(guido_image *img, double *out_ms, int32_t max_n, int32_t *n) {
    GUIDO_API_RANGE();
    if (!img || !img->dev || !out_ms || !n) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    DeviceState *ds = img->dev;
    long long have = std::min<long long>(ds->kev_count, DeviceState::kKev);
    long long take = std::min<long long>(have, max_n);
    for (long long i = 0; i < take; i++) {
        long long k = ds->kev_count - take + i;
        float ms = 0;
        GUIDO_CUDA(cudaEventElapsedTime(&ms, ds->kev[k % DeviceState::kKev][0], ds->kev[k % DeviceState::kKev][1]));
        out_ms[i] = ms;
    }
    *n = (int32_t)take;
    return GUIDO_OK;
}

int guido_image_upload(guido_image *img, int32_t device, int32_t shard_ix, int32_t n_shards) {
    GUIDO_API_RANGE();
    return upload_impl(img, device, shard_ix, n_shards);
}

int guido_image_get_info(const guido_image *img, guido_image_info *info) {
    GUIDO_API_RANGE();
    if (!img || !info) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    info->n_contigs = (int64_t)img->contigs.size();
    info->total_bases = img->total_bases;
    info->padded_bases = img->padded_bases;
    info->n_runs = (int64_t)img->runs.size();
    info->device = img->dev ? img->dev->device : -1;
    info->shard_lo = img->dev ? img->dev->shard_lo : 0;
    info->shard_hi = img->dev ? img->dev->shard_hi : img->padded_bases;
    info->reserved = 0;
    return GUIDO_OK;
}

int guido_pam_scan_genome(guido_image *img, const char *pam, int32_t mode, uint32_t *fw, int64_t cap_fw,
                          int64_t *n_fw, uint32_t *rv, int64_t cap_rv, int64_t *n_rv, guido_stats *stats) {
    GUIDO_API_RANGE();
    int rc = need_device(img);
    if (rc) return rc;
    PamSpec ps;
    if ((rc = parse_pam(pam, &ps))) return rc;
    if (mode < GUIDO_SCAN_REGEX || mode > GUIDO_SCAN_STRICT) { set_error("bad scan mode %d", mode); return GUIDO_ERR_ARG; }
    DeviceState *ds = img->dev;
    ScanRange r{(uint32_t)ds->shard_lo, (uint32_t)ds->shard_hi, (uint32_t)img->padded_bases, 0};
    return scan_to_host(img, ps, mode, r, fw, cap_fw, n_fw, rv, cap_rv, n_rv, stats);
}

int guido_pam_scan_genome_dev(guido_image *img, const char *pam, int32_t mode, void *d_fw, int64_t cap_fw,
                              void *d_rv, int64_t cap_rv, void *d_counts, void *stream) {
    GUIDO_API_RANGE();
    int rc = need_device(img);
    if (rc) return rc;
    PamSpec ps;
    if ((rc = parse_pam(pam, &ps))) return rc;
    if (mode < GUIDO_SCAN_REGEX || mode > GUIDO_SCAN_STRICT) { set_error("bad scan mode %d", mode); return GUIDO_ERR_ARG; }
    DeviceState *ds = img->dev;
    ScanRange r{(uint32_t)ds->shard_lo, (uint32_t)ds->shard_hi, (uint32_t)img->padded_bases, 0};
    return launch_pam_scan(ds, view_of(ds), ps, mode, r, (uint32_t *)d_fw, (uint64_t)cap_fw, (uint32_t *)d_rv,
                           (uint64_t)cap_rv, (unsigned long long *)d_counts, (cudaStream_t)stream, nullptr);
}

int guido_pam_scan_region(guido_image *img, int64_t contig_ix, int64_t start0, int64_t end0, const char *pam,
                          int32_t mode, int32_t *fw, int64_t cap_fw, int64_t *n_fw, int32_t *rv,
                          int64_t cap_rv, int64_t *n_rv, guido_stats *stats) {
    GUIDO_API_RANGE();
    int rc = need_device(img);
    if (rc) return rc;
    PamSpec ps;
    if ((rc = parse_pam(pam, &ps))) return rc;
    if (contig_ix < 0 || contig_ix >= (int64_t)img->contigs.size()) { set_error("contig index out of range"); return GUIDO_ERR_ARG; }
    const Contig &c = img->contigs[(size_t)contig_ix];
    if (start0 < 0 || end0 < start0 || end0 > c.len) { set_error("region outside contig"); return GUIDO_ERR_ARG; }
    DeviceState *ds = img->dev;
    int64_t lo = c.goff + start0, hi = c.goff + end0;
    if (lo < ds->shard_lo || hi > ds->shard_hi) { set_error("region is outside this handle's shard"); return GUIDO_ERR_ARG; }
    ScanRange r{(uint32_t)lo, (uint32_t)hi, (uint32_t)hi, (uint32_t)lo};
    return scan_to_host(img, ps, mode, r, (uint32_t *)fw, cap_fw, n_fw, (uint32_t *)rv, cap_rv, n_rv, stats);
}

int guido_pam_scan_seq(const char *seq, int64_t n, const char *pam, int32_t device, int32_t mode,
                       int32_t *fw, int64_t cap_fw, int64_t *n_fw, int32_t *rv, int64_t cap_rv,
                       int64_t *n_rv, guido_stats *stats) {
    GUIDO_API_RANGE();
    PamSpec ps;
    int rc = parse_pam(pam, &ps);
    if (rc) return rc;
    if (n < 0 || !seq) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    guido_image *img = nullptr;
    const char *name = "seq";
    if ((rc = guido_image_from_seqs(1, &name, &seq, &n, &img))) return rc;
    // The device side of this call -- stream, events, scratch, plane buffers -- is kept per device
    // between calls (a locus-sized scan takes microseconds; creating and destroying a context per
    // call took milliseconds): the cached DeviceState is lent to the temporary image.
    static std::mutex mu;
    static DeviceState *cached[64] = {};
    std::lock_guard<std::mutex> lock(mu);
    if (device >= 0 && device < 64 && cached[device]) { img->dev = cached[device]; cached[device] = nullptr; }
    rc = guido_image_upload(img, device, 0, 1);
    if (rc == GUIDO_OK)
        rc = guido_pam_scan_region(img, 0, 0, n, pam, mode, fw, cap_fw, n_fw, rv, cap_rv, n_rv, stats);
    if (stats && rc == GUIDO_OK) stats->bytes_h2d = img->dev->blk_n * 16;
    if (img->dev && device >= 0 && device < 64) { cached[device] = img->dev; img->dev = nullptr; }
    guido_image_free(img);
    return rc;
}

// ------------------------------------------------------------------------------- off-targets

static int make_ot_params(const char *pam, int core_len, int core_mm, int total_mm, bool raw, OtParams *op) {
    int rc = parse_pam(pam, &op->pam);
    if (rc) return rc;
    const int nN = op->pam.n_ambiguous;
    // In the bowtie read every non-ACGT PAM letter is an 'N' (helpers.py:24-26) and guido only
    // rejects mismatches at the A/C/G/T letters (off_targets.py:169-171), so for the off-target
    // site scan an IUPAC letter such as R constrains nothing: treat it as N.
    for (int i = 0; i < op->pam.len; i++)
        if (!(op->pam.fixed_mask & (1u << i))) op->pam.fw[i] = 15;
    for (int i = 0; i < op->pam.len; i++) {
        int s = op->pam.fw[op->pam.len - 1 - i];
        op->pam.rv[i] = (uint8_t)(((s & 1) << 3) | ((s & 2) << 1) | ((s & 4) >> 1) | ((s & 8) >> 3));
    }
    // same argument checks, same messages as guido/off_targets.py:141-149
    if (core_mm + nN > 3) {
        set_error("The value for the parameter core_mismatches is not valid: %d", core_mm);
        return GUIDO_ERR_ARG;
    }
    if (core_mm > total_mm) {
        set_error("The value for core_mismatches cannot be greater than total_mismatches: %d > %d", core_mm, total_mm);
        return GUIDO_ERR_ARG;
    }
    if (core_len < 0 || core_mm < 0) { set_error("negative core arguments"); return GUIDO_ERR_ARG; }
    // bowtie seed = first core_len + P read bases = PAM + PAM-proximal protospacer bases
    int cl = std::min(core_len, kProto);
    op->core_len = cl;
    uint32_t proto_mask = (1u << kProto) - 1;
    op->core_mask = cl ? (((1u << cl) - 1) << (kProto - cl)) : 0;
    op->cmp_mask = proto_mask;
    if (raw) {
        op->core_mask |= op->pam.fixed_mask << kProto;
        op->cmp_mask |= op->pam.fixed_mask << kProto;
    }
    op->core_max = core_mm;
    op->total_max = total_mm + 1 - nN;
    op->raw_mode = raw ? 1 : 0;
    return GUIDO_OK;
}

// Resolves GUIDO_OT_AUTO: the table join when the seed index applies and the PAM-site table exists
// or can be built (it is built here, on first use), else the streaming seed index, else brute force.
static int resolve_ot_algo(guido_image *img, const OtParams &op, const ScanRange &r, int64_t n_guides,
                           int32_t algo, cudaStream_t st, int32_t *out) {
    DeviceState *ds = img->dev;
    if (algo != GUIDO_OT_AUTO && algo != GUIDO_OT_SCAN && algo != GUIDO_OT_INDEX && algo != GUIDO_OT_TABLE) {
        set_error("unknown off-target algorithm %d", (int)algo);
        return GUIDO_ERR_ARG;
    }
    const int32_t algo_in = algo;
    if (algo == GUIDO_OT_AUTO) algo = ot_index_applicable(op, n_guides) ? GUIDO_OT_TABLE : GUIDO_OT_SCAN;
    else if (algo == GUIDO_OT_TABLE && !ot_index_applicable(op, n_guides)) algo = GUIDO_OT_INDEX;  // reports the reason
    if (algo == GUIDO_OT_TABLE && r.hi > r.lo && n_guides > 0) {
        const bool requested = algo_in == GUIDO_OT_TABLE;
        int rc = ensure_site_table(ds, view_of(ds), op, r, st, nullptr);
        if (rc == GUIDO_ERR_CAPACITY && !requested) {
            algo = GUIDO_OT_INDEX;  // AUTO: too many windows for HBM, stream the genome instead
        } else if (rc == GUIDO_ERR_CAPACITY) {
            // not GUIDO_ERR_CAPACITY: that code means "output buffer too small, call again"
            set_error("PAM-site table does not fit: more than 2^31 windows or not enough free device memory");
            return GUIDO_ERR_NOMEM;
        } else if (rc) {
            return rc;
        }
    }
    if (ds->site_n_parts > 1 && algo != GUIDO_OT_TABLE) {
        set_error("this handle covers 1/%d of the core-key space (guido_site_table_partition): only the "
                  "table join can search it", ds->site_n_parts);
        return GUIDO_ERR_ARG;
    }
    *out = algo;
    return GUIDO_OK;
}

int guido_site_table_partition(guido_image *img, int32_t part_ix, int32_t n_parts) {
    GUIDO_API_RANGE();
    int rc = need_device(img);
    if (rc) return rc;
    if (n_parts < 1 || n_parts > 32 || (n_parts & (n_parts - 1)) || part_ix < 0 || part_ix >= n_parts) {
        set_error("bad partition %d of %d: the number of parts must be a power of two <= 32", (int)part_ix, (int)n_parts);
        return GUIDO_ERR_ARG;
    }
    img->dev->site_part_ix = part_ix;  // the cached table (if any) no longer matches and is rebuilt on demand
    img->dev->site_n_parts = n_parts;
    return GUIDO_OK;
}

int guido_site_table_build(guido_image *img, const char *pam, int32_t core_len, int64_t *n_sites, double *build_ms) {
    GUIDO_API_RANGE();
    int rc = need_device(img);
    if (rc) return rc;
    OtParams op;
    if ((rc = make_ot_params(pam, core_len, 0, 0, false, &op))) return rc;
    DeviceState *ds = img->dev;
    ScanRange r{(uint32_t)ds->shard_lo, (uint32_t)ds->shard_hi, (uint32_t)img->padded_bases, 0};
    bool built = false;
    rc = ensure_site_table(ds, view_of(ds), op, r, ds->stream, &built);
    if (rc == GUIDO_ERR_CAPACITY) {
        set_error("PAM-site table does not fit: core_length outside 6..12, more than 2^31 windows or not enough free device memory");
        return GUIDO_ERR_NOMEM;
    }
    if (rc) return rc;
    if (n_sites) *n_sites = ds->site_n;
    if (build_ms) *build_ms = built ? (double)ds->site_build_ms : 0.0;
    return GUIDO_OK;
}

int guido_site_table_free(guido_image *img) {
    GUIDO_API_RANGE();
    if (img && img->dev) release_site_table(img->dev);
    return GUIDO_OK;
}

int guido_offtarget_search_dev(guido_image *img, const void *d_guides, int64_t n_guides, const char *pam,
                               int32_t core_len, int32_t core_mm, int32_t total_mm, int32_t algo,
                               void *d_hits, int64_t cap, void *d_count, void *stream) {
    GUIDO_API_RANGE();
    int rc = need_device(img);
    if (rc) return rc;
    OtParams op;
    if ((rc = make_ot_params(pam, core_len, core_mm, total_mm, false, &op))) return rc;
    DeviceState *ds = img->dev;
    ScanRange r{(uint32_t)ds->shard_lo, (uint32_t)ds->shard_hi, (uint32_t)img->padded_bases, 0};
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = resolve_ot_algo(img, op, r, n_guides, algo, st, &algo))) return rc;
    GUIDO_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st));
    if (algo == GUIDO_OT_TABLE)
        return launch_ot_join(ds, view_of(ds), op, r, (const uint2 *)d_guides, n_guides, (guido_ot_hit *)d_hits,
                              (uint64_t)cap, (unsigned long long *)d_count, nullptr, st, nullptr);
    if (algo == GUIDO_OT_INDEX)
        return launch_ot_index(ds, view_of(ds), op, r, (const uint2 *)d_guides, n_guides, (guido_ot_hit *)d_hits,
                               (uint64_t)cap, (unsigned long long *)d_count, nullptr, st, nullptr);
    return launch_ot_scan(ds, view_of(ds), op, r, (const uint2 *)d_guides, n_guides, (guido_ot_hit *)d_hits,
                          (uint64_t)cap, (unsigned long long *)d_count, nullptr, nullptr, st, nullptr);
}

int guido_offtarget_search_sliced_dev(guido_image *img, const void *d_guides, int64_t n_guides, const char *pam,
                                      int32_t core_len, int32_t core_mm, int32_t total_mm, int32_t n_slices,
                                      void *d_hits, int64_t cap_per_slice, void *d_counts, void *const *events,
                                      void *d_send, int64_t send_stride, void *stream) {
    GUIDO_API_RANGE();
    int rc = need_device(img);
    if (rc) return rc;
    if (n_slices < 1 || n_slices > 8 || (n_slices & (n_slices - 1)) || cap_per_slice < 0 || !d_counts) {
        set_error("n_slices must be 1, 2, 4 or 8"); return GUIDO_ERR_ARG;
    }
    OtParams op;
    if ((rc = make_ot_params(pam, core_len, core_mm, total_mm, false, &op))) return rc;
    DeviceState *ds = img->dev;
    ScanRange r{(uint32_t)ds->shard_lo, (uint32_t)ds->shard_hi, (uint32_t)img->padded_bases, 0};
    cudaStream_t st = (cudaStream_t)stream;
    int algo = GUIDO_OT_TABLE;
    if ((rc = resolve_ot_algo(img, op, r, n_guides, algo, st, &algo))) return rc;
    if (algo != GUIDO_OT_TABLE) { set_error("the sliced search needs the PAM-site table"); return GUIDO_ERR_ARG; }
    GUIDO_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)n_slices * sizeof(unsigned long long), st));
    OtSlices sl;
    cudaEvent_t ev[8] = {};
    sl.n = n_slices;
    if (events) { for (int s = 0; s < n_slices; s++) ev[s] = (cudaEvent_t)events[s]; sl.events = ev; }
    if (d_send) {
        if (send_stride < 2) { set_error("send_stride must hold the header word and at least one record"); return GUIDO_ERR_ARG; }
        sl.send = (unsigned long long *)d_send; sl.send_stride = (uint64_t)send_stride;
    }
    return launch_ot_join(ds, view_of(ds), op, r, (const uint2 *)d_guides, n_guides, (guido_ot_hit *)d_hits,
                          (uint64_t)cap_per_slice, (unsigned long long *)d_counts, nullptr, st, nullptr, &sl);
}

int guido_encode_guides_dev(guido_image *img, const void *d_letters, int64_t n_guides, void *d_out, void *d_bad,
                            void *stream) {
    GUIDO_API_RANGE();
    int rc = need_device(img);
    if (rc) return rc;
    if (n_guides < 0 || n_guides > 0x7fffffff) { set_error("bad guide count"); return GUIDO_ERR_ARG; }
    DeviceState *ds = img->dev;
    // without a report slot bad letters go unnoticed (the handle's own slot takes the atomics)
    unsigned long long *bad = d_bad ? (unsigned long long *)d_bad : ds->d_counts + 6;
    return encode_guides_device(ds, (const char *)d_letters, n_guides, (uint2 *)d_out, bad, (cudaStream_t)stream);
}

int guido_hits_concat_dev(guido_image *img, const void *d_gathered, int32_t n_parts, int64_t stride, void *d_out,
                          int64_t cap, void *d_total, void *stream) {
    GUIDO_API_RANGE();
    int rc = need_device(img);
    if (rc) return rc;
    if (!d_gathered || !d_out || !d_total || stride < 1 || cap < 0) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    return concat_hits_device(img->dev, (const guido_ot_hit *)d_gathered, n_parts, (uint64_t)stride, (guido_ot_hit *)d_out,
                              (uint64_t)cap, (unsigned long long *)d_total, (cudaStream_t)stream);
}

int guido_hits_pack_dev(guido_image *img, const void *d_hits, int64_t n_hits, void *d_out, void *stream) {
    GUIDO_API_RANGE();
    int rc = need_device(img);
    if (rc) return rc;
    if (n_hits < 0 || (n_hits && (!d_hits || !d_out))) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    return pack_hits_device(img->dev, (const guido_ot_hit *)d_hits, (uint64_t)n_hits, (unsigned long long *)d_out, (cudaStream_t)stream);
}

int guido_hits_pack_send_dev(guido_image *img, const void *d_hits, const void *d_count, int64_t cap, void *d_send, void *stream) {
    GUIDO_API_RANGE();
    int rc = need_device(img);
    if (rc) return rc;
    if (!d_hits || !d_count || !d_send || cap < 0) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    return pack_send_device(img->dev, (const guido_ot_hit *)d_hits, (const unsigned long long *)d_count, (uint64_t)cap,
                            (unsigned long long *)d_send, (cudaStream_t)stream);
}

int guido_packed_concat_dev(guido_image *img, const void *d_gathered, int32_t n_parts, int64_t stride, void *d_out,
                            int64_t cap, void *d_total, void *stream) {
    GUIDO_API_RANGE();
    int rc = need_device(img);
    if (rc) return rc;
    if (!d_gathered || !d_out || !d_total || stride < 1 || cap < 0) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    return concat_packed_device(img->dev, (const unsigned long long *)d_gathered, n_parts, (uint64_t)stride,
                                (unsigned long long *)d_out, (uint64_t)cap, (unsigned long long *)d_total, (cudaStream_t)stream);
}

int guido_hits_sort_dev(guido_image *img, void *d_hits, int64_t n_hits, int64_t n_guides, void *stream) {
    GUIDO_API_RANGE();
    int rc = need_device(img);
    if (rc) return rc;
    if (n_hits < 0 || n_guides < 1) { set_error("bad hit or guide count"); return GUIDO_ERR_ARG; }
    return sort_hits_device(img->dev, (guido_ot_hit *)d_hits, (uint64_t)n_hits, n_guides, (cudaStream_t)stream, nullptr);
}

// Key presence (off_targets.py:183-184) for the guides listed in `todo`: does the guide have ANY raw
// bowtie alignment in this handle's stretch of the genome -- PAM mismatches allowed inside the seed
// budget?  Brute force over every window (the PAM's fixed letters join the compare), flags[todo[i]].
static int relaxed_flag_pass(guido_image *img, const char *protos, const std::vector<int64_t> &todo, const char *pam,
                             int core_len, int core_mm, int total_mm, uint8_t *flags, float *kernel_ms,
                             int *launches, int64_t *h2d, int64_t *d2h) {
    if (todo.empty()) return GUIDO_OK;
    DeviceState *ds = img->dev;
    cudaStream_t st = ds->stream;
    ScanRange r{(uint32_t)ds->shard_lo, (uint32_t)ds->shard_hi, (uint32_t)img->padded_bases, 0};
    OtParams rp;
    int rc;
    if ((rc = make_ot_params(pam, core_len, core_mm, total_mm, true, &rp))) return rc;
    std::vector<uint32_t> sub(2 * todo.size());
    for (size_t i = 0; i < todo.size(); i++) {
        uint32_t hl[2];
        if ((rc = guido_encode_guides(protos + kProto * todo[i], 1, hl))) return rc;
        uint32_t hi = hl[0], lo = hl[1];
        for (int k = 0; k < rp.pam.len; k++)
            if (rp.pam.fixed_mask & (1u << k)) {
                hi |= (uint32_t)(rp.pam.code[k] >> 1) << (kProto + k);
                lo |= (uint32_t)(rp.pam.code[k] & 1) << (kProto + k);
            }
        sub[2 * i] = hi; sub[2 * i + 1] = lo;
    }
    if ((rc = ensure_bytes(&ds->d_flags, &ds->flags_cap, (int64_t)todo.size()))) return rc;
    if ((rc = ensure_bytes(&ds->d_guides, &ds->guides_cap, (int64_t)sub.size() * 4))) return rc;
    GUIDO_CUDA(cudaMemsetAsync(ds->d_flags, 0, todo.size(), st));
    GUIDO_CUDA(cudaMemcpyAsync(ds->d_guides, sub.data(), sub.size() * 4, cudaMemcpyHostToDevice, st));
    GUIDO_CUDA(cudaEventRecord(ds->ev[1], st));
    rc = launch_ot_scan(ds, view_of(ds), rp, r, (const uint2 *)ds->d_guides, (int64_t)todo.size(), nullptr, 0,
                        ds->d_counts, (uint8_t *)ds->d_flags, nullptr, st, launches);
    if (rc) return rc;
    GUIDO_CUDA(cudaEventRecord(ds->ev[2], st));
    std::vector<uint8_t> fl(todo.size());
    GUIDO_CUDA(cudaMemcpyAsync(fl.data(), ds->d_flags, todo.size(), cudaMemcpyDeviceToHost, st));
    GUIDO_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, ds->ev[1], ds->ev[2]);
    if (kernel_ms) *kernel_ms += ms;
    for (size_t i = 0; i < todo.size(); i++) flags[todo[i]] = fl[i];
    if (h2d) *h2d += (int64_t)sub.size() * 4;
    if (d2h) *d2h += (int64_t)todo.size();
    return GUIDO_OK;
}

int guido_offtarget_raw_flags(guido_image *img, const char *protos, int64_t n_guides, const char *pam,
                              int32_t core_len, int32_t core_mm, int32_t total_mm, uint8_t *flags) {
    GUIDO_API_RANGE();
    int rc = need_device(img);
    if (rc) return rc;
    if (n_guides < 0 || n_guides > 0x7fffffff) { set_error("bad guide count"); return GUIDO_ERR_ARG; }
    std::vector<int64_t> todo((size_t)n_guides);
    for (int64_t g = 0; g < n_guides; g++) todo[(size_t)g] = g;
    memset(flags, 0, (size_t)n_guides);
    return relaxed_flag_pass(img, protos, todo, pam, core_len, core_mm, total_mm, flags, nullptr, nullptr, nullptr, nullptr);
}

}  // extern "C"

// what goes to the host: guido_ot_hit records, 8-byte records (guide << 33 | gpos << 1 | strand), or CSR by
// guide (hits_out = positions, plus strand bits and per-guide offsets)
enum HitFormat { kHitsFull = 0, kHitsPacked = 1, kHitsCsr = 2 };

static int offtarget_search_impl(guido_image *img, const char *protos, int64_t n_guides, const char *pam,
                                 int32_t core_len, int32_t core_mm, int32_t total_mm, int32_t algo,
                                 void *hits_out, HitFormat fmt, int64_t cap, int64_t *n_hits, uint8_t *raw_flags,
                                 guido_stats *stats, uint32_t *csr_strand = nullptr, int64_t *csr_offsets = nullptr) {
    const bool packed = fmt == kHitsPacked;
    int rc = need_device(img);
    if (rc) return rc;
    OtParams op;
    if ((rc = make_ot_params(pam, core_len, core_mm, total_mm, false, &op))) return rc;
    if (n_guides < 0 || n_guides > 0x7fffffff) { set_error("bad guide count"); return GUIDO_ERR_ARG; }
    DeviceState *ds = img->dev;
    cudaStream_t st = ds->stream;
    // GUIDO_TRACE=1: host wall-clock of the call's phases on stderr (tools/e2e_probe.py)
    const bool trace = getenv("GUIDO_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (trace) fprintf(stderr, "[guido] %-18s %8.3f ms\n", what,
                           std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    };
    // the letters go to the device as they are and are packed into plane words there (a host loop
    // over 2e6 letters costs more than the whole search); d_guides = [n x uint2][n x 20 letters]
    if ((rc = ensure_bytes(&ds->d_guides, &ds->guides_cap, std::max<int64_t>(28 * n_guides, 8) + 64))) return rc;
    ScanRange r{(uint32_t)ds->shard_lo, (uint32_t)ds->shard_hi, (uint32_t)img->padded_bases, 0};
    if ((rc = resolve_ot_algo(img, op, r, n_guides, algo, st, &algo))) return rc;
    const bool use_index = algo != GUIDO_OT_SCAN;
    int launches = 0;
    GUIDO_CUDA(cudaEventRecord(ds->ev[0], st));
    {
        char *d_letters = (char *)ds->d_guides + 8 * n_guides;
        GUIDO_CUDA(cudaMemcpyAsync(d_letters, protos, (size_t)(kProto * n_guides), cudaMemcpyHostToDevice, st));
        GUIDO_CUDA(cudaMemsetAsync(ds->d_counts + 6, 0xff, sizeof(unsigned long long), st));
        if ((rc = encode_guides_device(ds, d_letters, n_guides, (uint2 *)ds->d_guides, ds->d_counts + 6, st))) return rc;
        unsigned long long bad = 0;
        GUIDO_CUDA(cudaMemcpyAsync(&bad, ds->d_counts + 6, sizeof bad, cudaMemcpyDeviceToHost, st));
        GUIDO_CUDA(cudaStreamSynchronize(st));
        if (bad != ~0ull) {
            set_error("guide %lld holds a letter outside A/C/G/T at position %d", (long long)(bad / 32), (int)(bad % 32) + 1);
            return GUIDO_ERR_ARG;
        }
        launches += 1;
    }
    lap("guides on device");
    // device-side capacity: the caller's, a per-guide guess, or what the previous search needed
    int64_t dcap = std::max<int64_t>(std::max<int64_t>(cap, 64 * n_guides), std::max<int64_t>(1 << 16, ds->hits_hint + ds->hits_hint / 16));
    unsigned long long found = 0, sites = 0, visited = 0;
    float kernel_ms = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        if ((rc = ensure_bytes(&ds->d_out[0], &ds->out_cap[0], dcap * (int64_t)sizeof(guido_ot_hit)))) return rc;
        GUIDO_CUDA(cudaMemsetAsync(ds->d_counts, 0, 4 * sizeof(unsigned long long), st));
        GUIDO_CUDA(cudaEventRecord(ds->ev[1], st));
        if (algo == GUIDO_OT_TABLE)
            rc = launch_ot_join(ds, view_of(ds), op, r, (const uint2 *)ds->d_guides, n_guides,
                                (guido_ot_hit *)ds->d_out[0], (uint64_t)dcap, ds->d_counts, ds->d_counts + 1, st, &launches);
        else if (use_index)
            rc = launch_ot_index(ds, view_of(ds), op, r, (const uint2 *)ds->d_guides, n_guides,
                                 (guido_ot_hit *)ds->d_out[0], (uint64_t)dcap, ds->d_counts, ds->d_counts + 1, st, &launches);
        else
            rc = launch_ot_scan(ds, view_of(ds), op, r, (const uint2 *)ds->d_guides, n_guides,
                                (guido_ot_hit *)ds->d_out[0], (uint64_t)dcap, ds->d_counts, nullptr,
                                ds->d_counts + 1, st, &launches);
        if (rc) return rc;
        GUIDO_CUDA(cudaEventRecord(ds->ev[2], st));
        unsigned long long c2[4];
        GUIDO_CUDA(cudaMemcpyAsync(c2, ds->d_counts, sizeof c2, cudaMemcpyDeviceToHost, st));
        GUIDO_CUDA(cudaStreamSynchronize(st));
        if (c2[3]) { set_error("internal: %llu seed-index buckets were gathered with a wrong size", c2[3]); return GUIDO_ERR_CUDA; }
        float ms = 0;
        cudaEventElapsedTime(&ms, ds->ev[1], ds->ev[2]);
        kernel_ms += ms;
        found = c2[0]; sites = c2[1]; visited = c2[2];
        lap("search done");
        if ((int64_t)found <= dcap) break;
        dcap = (int64_t)found;  // exact size known now: one more pass
    }
    *n_hits = (int64_t)found;
    ds->hits_hint = (int64_t)found;
    // canonical (guide, gpos, strand) order on the device, then one copy straight into the
    // caller's buffer (fast when that buffer is pinned, see guido_host_alloc)
    if ((rc = sort_hits_device(ds, (guido_ot_hit *)ds->d_out[0], found, n_guides, st, &launches))) return rc;
    if (trace) { cudaStreamSynchronize(st); lap("sort done"); }
    int64_t ncopy = std::min<int64_t>((int64_t)found, std::max<int64_t>(cap, 0));
    const int64_t rec_bytes = packed ? 8 : fmt == kHitsCsr ? 4 : (int64_t)sizeof(guido_ot_hit);
    int64_t csr_extra = 0;
    if (fmt == kHitsCsr) {
        // positions, strand bits and per-guide offsets are cut out of the sorted records on the device
        // (the sort's scratch is free again); the offsets travel as 32-bit words and are widened here
        const int64_t n_words = (ncopy + 31) / 32, n_off = n_guides + 1;
        if ((rc = ensure_bytes(&ds->d_out[1], &ds->out_cap[1], (ncopy + n_words + n_off + 16) * 4))) return rc;
        if ((rc = ensure_pinned(ds, n_off * 4))) return rc;
        uint32_t *d_gpos = (uint32_t *)ds->d_out[1], *d_bits = d_gpos + ncopy, *d_off = d_bits + n_words;
        if ((rc = csr_hits_device(ds, (const guido_ot_hit *)ds->d_out[0], (uint64_t)ncopy, (uint32_t)n_guides, d_gpos, d_bits, d_off, st))) return rc;
        launches += 1;
        if (ncopy) {
            GUIDO_CUDA(cudaMemcpyAsync(hits_out, d_gpos, (size_t)ncopy * 4, cudaMemcpyDeviceToHost, st));
            GUIDO_CUDA(cudaMemcpyAsync(csr_strand, d_bits, (size_t)n_words * 4, cudaMemcpyDeviceToHost, st));
        }
        GUIDO_CUDA(cudaMemcpyAsync(ds->h_pinned, d_off, (size_t)n_off * 4, cudaMemcpyDeviceToHost, st));
        csr_extra = n_words * 4 + n_off * 4;
    } else if (ncopy && packed) {
        // the sort's scratch (d_out[1]) is free again: pack into it
        if ((rc = ensure_bytes(&ds->d_out[1], &ds->out_cap[1], ncopy * 8))) return rc;
        if ((rc = pack_hits_device(ds, (const guido_ot_hit *)ds->d_out[0], (uint64_t)ncopy, (unsigned long long *)ds->d_out[1], st))) return rc;
        launches += 1;
        GUIDO_CUDA(cudaMemcpyAsync(hits_out, ds->d_out[1], (size_t)ncopy * 8, cudaMemcpyDeviceToHost, st));
    } else if (ncopy) {
        GUIDO_CUDA(cudaMemcpyAsync(hits_out, ds->d_out[0], (size_t)ncopy * sizeof(guido_ot_hit), cudaMemcpyDeviceToHost, st));
    }
    int64_t d2h = 16 + ncopy * rec_bytes + csr_extra, h2d = kProto * n_guides;
    if (raw_flags && n_guides > 0) {
        // a valid hit is a raw alignment: mark the guides that own one (on the device -- walking
        // 1e7 hit records on the host costs more than the search)
        if ((rc = ensure_bytes(&ds->d_flags, &ds->flags_cap, n_guides))) return rc;
        GUIDO_CUDA(cudaMemsetAsync(ds->d_flags, 0, (size_t)n_guides, st));
        if ((rc = mark_hit_guides(ds, (const guido_ot_hit *)ds->d_out[0], std::min<uint64_t>(found, (uint64_t)dcap),
                                  (uint8_t *)ds->d_flags, st))) return rc;
        GUIDO_CUDA(cudaMemcpyAsync(raw_flags, ds->d_flags, (size_t)n_guides, cudaMemcpyDeviceToHost, st));
        d2h += n_guides;
        launches += 1;
    }
    GUIDO_CUDA(cudaEventRecord(ds->ev[3], st));
    GUIDO_CUDA(cudaStreamSynchronize(st));
    if (fmt == kHitsCsr) {
        const uint32_t *off32 = (const uint32_t *)ds->h_pinned;
        for (int64_t g = 0; g <= n_guides; g++) csr_offsets[g] = off32[g];
    }
    lap("copy done");
    if ((int64_t)found > cap) {
        set_error("output capacity too small: need %lld hits", (long long)found);
        return GUIDO_ERR_CAPACITY;
    }

    // On a handle that covers 1/n of the core-key space the flags stay "has a valid hit in this
    // part": most guides have none in a given part, and the relaxed pass below would brute-force
    // all of them against the whole genome on every rank.  The caller ORs the flags of the parts and
    // asks an unpartitioned handle about the guides that are left (guido_b200/parallel.py).
    if (raw_flags && ds->site_n_parts == 1) {
        // key presence (off_targets.py:183-184): any raw bowtie alignment, PAM mismatches allowed
        // inside the seed budget.  Only guides without a valid hit need the relaxed all-window pass.
        std::vector<int64_t> todo;
        for (int64_t g = 0; g < n_guides; g++) if (!raw_flags[g]) todo.push_back(g);
        if ((rc = relaxed_flag_pass(img, protos, todo, pam, core_len, core_mm, total_mm, raw_flags, &kernel_ms,
                                    &launches, &h2d, &d2h))) return rc;
    }
    if (stats) {
        float t = 0;
        cudaEventElapsedTime(&t, ds->ev[0], ds->ev[3]);
        stats->kernel_ms = kernel_ms; stats->total_ms = t; stats->n_launches = launches;
        stats->n_sites = (int64_t)sites;
        stats->n_pairs = use_index ? (int64_t)visited : (int64_t)sites * n_guides;
        stats->bytes_h2d = h2d; stats->bytes_d2h = d2h;
    }
    if ((int64_t)found > cap) {
        set_error("output capacity too small: need %lld hits", (long long)found);
        return GUIDO_ERR_CAPACITY;
    }
    return GUIDO_OK;
}

extern "C" {

int guido_offtarget_search(guido_image *img, const char *protos, int64_t n_guides, const char *pam,
                           int32_t core_len, int32_t core_mm, int32_t total_mm, int32_t algo,
                           guido_ot_hit *hits, int64_t cap, int64_t *n_hits, uint8_t *raw_flags,
                           guido_stats *stats) {
    GUIDO_API_RANGE();
    return offtarget_search_impl(img, protos, n_guides, pam, core_len, core_mm, total_mm, algo, hits, kHitsFull, cap, n_hits,
                                 raw_flags, stats);
}

int guido_offtarget_search_packed(guido_image *img, const char *protos, int64_t n_guides, const char *pam,
                                  int32_t core_len, int32_t core_mm, int32_t total_mm, int32_t algo,
                                  uint64_t *hits, int64_t cap, int64_t *n_hits, uint8_t *raw_flags,
                                  guido_stats *stats) {
    GUIDO_API_RANGE();
    return offtarget_search_impl(img, protos, n_guides, pam, core_len, core_mm, total_mm, algo, hits, kHitsPacked, cap, n_hits,
                                 raw_flags, stats);
}

int guido_offtarget_search_csr(guido_image *img, const char *protos, int64_t n_guides, const char *pam,
                               int32_t core_len, int32_t core_mm, int32_t total_mm, int32_t algo,
                               uint32_t *gpos, uint32_t *strand_bits, int64_t cap, int64_t *guide_offsets, int64_t *n_hits,
                               uint8_t *raw_flags, guido_stats *stats) {
    GUIDO_API_RANGE();
    if (!guide_offsets || (cap > 0 && (!gpos || !strand_bits))) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    return offtarget_search_impl(img, protos, n_guides, pam, core_len, core_mm, total_mm, algo, gpos, kHitsCsr, cap, n_hits,
                                 raw_flags, stats, strand_bits, guide_offsets);
}

}  // extern "C"
