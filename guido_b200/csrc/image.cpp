// image.cpp -- host side of the genome image: FASTA -> bit planes + non-ACGT runs + contig table.
//
// Replaces the `bowtie-build` step of Genome.build (reference guido/genome.py:156-171) for the
// accelerated path.  Layout of the global coordinate:
//
//     [pad 64][contig 0 ....][pad >=64, next start 64-aligned][contig 1 ....] ... [pad >=64]
//
// Every pad is recorded as a kind-0 run, so "window touches a run" covers N bases, contig ends
// (bowtie: alignments may not fall off a reference sequence) and the genome ends in one test.
// Bases are stored as two bit planes, 64 bases per {lo,hi} word pair: lo = code&1, hi = code>>1
// with A=0 C=1 G=2 T=3; bases inside runs are stored as code 0.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdarg>
#include <cstring>

#include <nvtx3/nvToolsExt.h>

#include "internal.h"

namespace guido {

ApiRange::ApiRange(const char *name) { nvtxRangePushA(name); }
ApiRange::~ApiRange() { nvtxRangePop(); }


static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

static int iupac_set(char c) {
    switch (c) {  // reference guido/locus.py:160-176
        case 'A': return 1;
        case 'C': return 2;
        case 'G': return 4;
        case 'T': return 8;
        case 'R': return 1 | 4;
        case 'Y': return 2 | 8;
        case 'S': return 2 | 4;
        case 'W': return 1 | 8;
        case 'K': return 4 | 8;
        case 'M': return 1 | 2;
        case 'B': return 2 | 4 | 8;
        case 'D': return 1 | 4 | 8;
        case 'H': return 1 | 2 | 8;
        case 'V': return 1 | 2 | 4;
        case 'N': return 15;
        default: return 0;
    }
}

int parse_pam(const char *pam, PamSpec *spec) {
    memset(spec, 0, sizeof *spec);
    int P = pam ? (int)strlen(pam) : 0;
    if (P < 1 || P > kMaxPam) {
        set_error("PAM length must be 1..%d, got %d", kMaxPam, P);
        return GUIDO_ERR_ARG;
    }
    spec->len = P;
    for (int i = 0; i < P; i++) {
        int s = iupac_set(pam[i]);
        if (!s) {
            set_error("'%c'", pam[i]);  // KeyError text of the reference's dict lookup
            return GUIDO_ERR_PAM;
        }
        spec->fw[i] = (uint8_t)s;
        if (s == 1 || s == 2 || s == 4 || s == 8) {
            spec->fixed_mask |= 1u << i;
            spec->code[i] = (uint8_t)(s == 1 ? 0 : s == 2 ? 1 : s == 4 ? 2 : 3);
        }
        // rev_comp() of the reference keeps A/C/G/T/N and maps every other letter to N
        // (helpers.py:19-26), and run_bowtie counts the N's of rev_comp(pam) (off_targets.py:128)
        if (!(spec->fixed_mask & (1u << i))) spec->n_ambiguous++;
    }
    for (int i = 0; i < P; i++) {
        int s = spec->fw[P - 1 - i];
        spec->rv[i] = (uint8_t)(((s & 1) << 3) | ((s & 2) << 1) | ((s & 4) >> 1) | ((s & 8) >> 3));
    }
    return GUIDO_OK;
}

// ------------------------------------------------------------------------------------- packer

struct Packer {
    guido_image *img;
    int64_t g;       // next global position
    uint64_t lo = 0, hi = 0;
    bool have_run = false;
    Run run{};
    static const int8_t *lut() {
        static int8_t t[256];
        static bool init = false;
        if (!init) {
            for (int i = 0; i < 256; i++) t[i] = 4;
            t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3;
            t['\n'] = t['\r'] = 5;
            init = true;
        }
        return t;
    }
    Packer(guido_image *im, int64_t start) : img(im), g(start) {}
    inline void flush_word() {
        int64_t b = (g - 1) / kBlockBases;
        img->planes[2 * b] = lo;
        img->planes[2 * b + 1] = hi;
        lo = hi = 0;
    }
    inline void put_code(int c) {
        int s = (int)(g & 63);
        lo |= (uint64_t)(c & 1) << s;
        hi |= (uint64_t)(c >> 1) << s;
        g++;
        if ((g & 63) == 0) flush_word();
    }
    void end_run() {
        if (have_run) { img->runs.push_back(run); have_run = false; }
    }
    void put_other(char letter) {
        if (letter >= 'a' && letter <= 'z') letter = (char)(letter - 32);
        if (have_run && run.letter == letter && run.end == (uint32_t)g) {
            run.end++;
        } else {
            end_run();
            run.start = (uint32_t)g; run.end = (uint32_t)g + 1;
            run.letter = letter; run.kind = (letter == 'N') ? 0 : 1;
            have_run = true;
        }
        put_code(0);
    }
    void push(const char *s, int64_t n, bool skip_newlines) {
        const int8_t *t = lut();
        for (int64_t i = 0; i < n; i++) {
            int c = t[(unsigned char)s[i]];
            if (c < 4) put_code(c);
            else if (c == 5 && skip_newlines) continue;
            else put_other(s[i]);
        }
    }
    void finish() {
        end_run();
        if (g & 63) { g = (g + 63) & ~(int64_t)63; flush_word(); }
    }
};

static inline int64_t round_up64(int64_t x) { return (x + 63) & ~(int64_t)63; }

int image_finalize_layout(guido_image *img) {
    int64_t g = kBlockBases;
    img->total_bases = 0;
    for (auto &c : img->contigs) {
        c.goff = g;
        g = round_up64(g + c.len) + kBlockBases;
        img->total_bases += c.len;
    }
    img->padded_bases = g;
    if (img->padded_bases >= (int64_t)0xFFFFFF00ll) {
        set_error("genome too large for 32-bit global positions (%lld bases)", (long long)g);
        return GUIDO_ERR_ARG;
    }
    img->n_blocks = g / kBlockBases;
    img->planes.assign((size_t)(2 * img->n_blocks), 0);
    return GUIDO_OK;
}

void image_add_pads(guido_image *img) {
    // pads: [0, goff0), [end_c, goff_{c+1}), [end_last, padded)
    std::vector<Run> pads;
    int64_t prev_end = 0;
    for (auto &c : img->contigs) {
        pads.push_back(Run{(uint32_t)prev_end, (uint32_t)c.goff, 0, 0});
        prev_end = c.goff + c.len;
    }
    pads.push_back(Run{(uint32_t)prev_end, (uint32_t)img->padded_bases, 0, 0});
    std::vector<Run> all;
    all.reserve(img->runs.size() + pads.size());
    std::merge(img->runs.begin(), img->runs.end(), pads.begin(), pads.end(), std::back_inserter(all),
               [](const Run &a, const Run &b) { return a.start < b.start; });
    img->runs.swap(all);
}

void image_pack_contig(guido_image *img, size_t ci, const char *seq, int64_t len) {
    Packer p(img, img->contigs[ci].goff);
    p.push(seq, len, false);
    p.finish();
}

int64_t image_find_contig(const guido_image *img, int64_t gpos) {
    int64_t lo = 0, hi = (int64_t)img->contigs.size() - 1, ans = -1;
    while (lo <= hi) {
        int64_t mid = (lo + hi) / 2;
        if (img->contigs[mid].goff <= gpos) { ans = mid; lo = mid + 1; } else hi = mid - 1;
    }
    if (ans < 0) return -1;
    if (gpos >= img->contigs[ans].goff + img->contigs[ans].len) return -1;
    return ans;
}

}  // namespace guido

using namespace guido;

extern "C" {

const char *guido_last_error(void) { return g_err; }
int guido_version(void) { return 2; }

int guido_image_from_seqs(int32_t n, const char *const *names, const char *const *seqs,
                          const int64_t *lens, guido_image **out) {
    GUIDO_API_RANGE();
    if (n < 0 || !out) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    auto *img = new guido_image();
    for (int i = 0; i < n; i++) img->contigs.push_back(Contig{names ? names[i] : "seq", lens[i], 0});
    int rc = image_finalize_layout(img);
    if (rc) { delete img; return rc; }
    for (int i = 0; i < n; i++) image_pack_contig(img, (size_t)i, seqs[i], lens[i]);
    image_add_pads(img);
    *out = img;
    return GUIDO_OK;
}

int guido_image_from_fasta(const char *path, guido_image **out) {
    GUIDO_API_RANGE();
    int fd = open(path, O_RDONLY);
    if (fd < 0) { set_error("cannot open %s", path); return GUIDO_ERR_IO; }
    struct stat st;
    fstat(fd, &st);
    size_t sz = (size_t)st.st_size;
    const char *buf = sz ? (const char *)mmap(nullptr, sz, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
    if (sz && buf == MAP_FAILED) { close(fd); set_error("mmap failed for %s", path); return GUIDO_ERR_IO; }
    auto *img = new guido_image();
    // pass 1: names, sequence byte ranges, lengths
    struct Rec { size_t b, e; };
    std::vector<Rec> recs;
    size_t i = 0;
    while (i < sz) {
        if (buf[i] == '>') {
            size_t j = i + 1;
            while (j < sz && buf[j] != '\n') j++;
            size_t k = i + 1;
            while (k < j && buf[k] != ' ' && buf[k] != '\t' && buf[k] != '\r') k++;
            img->contigs.push_back(Contig{std::string(buf + i + 1, k - i - 1), 0, 0});
            size_t b = (j < sz) ? j + 1 : sz, e = b;
            int64_t len = 0;
            while (e < sz && buf[e] != '>') {
                const char *nl = (const char *)memchr(buf + e, '\n', sz - e);
                size_t le = nl ? (size_t)(nl - buf) : sz;
                size_t ll = le - e;
                // carriage returns are skipped wherever they stand, exactly as the packing pass does
                if (ll && memchr(buf + e, '\r', ll)) ll -= (size_t)std::count(buf + e, buf + le, '\r');
                len += (int64_t)ll;
                e = nl ? le + 1 : sz;
            }
            img->contigs.back().len = len;
            recs.push_back(Rec{b, e});
            i = e;
        } else {
            const char *nl = (const char *)memchr(buf + i, '\n', sz - i);
            i = nl ? (size_t)(nl - buf) + 1 : sz;
        }
    }
    int rc = image_finalize_layout(img);
    if (rc == GUIDO_OK) {
        for (size_t c = 0; c < recs.size(); c++) {
            Packer p(img, img->contigs[c].goff);
            p.push(buf + recs[c].b, (int64_t)(recs[c].e - recs[c].b), true);
            p.finish();
        }
        image_add_pads(img);
    }
    if (sz) munmap((void *)buf, sz);
    close(fd);
    if (rc) { delete img; return rc; }
    *out = img;
    return GUIDO_OK;
}

int guido_image_synth(int32_t n, const char *const *names, const int64_t *lens, uint64_t seed,
                      double n_frac, int64_t run_min, int64_t run_max, guido_image **out) {
    GUIDO_API_RANGE();
    if (n < 0 || !out || run_min < 1 || run_max < run_min) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    auto *img = new guido_image();
    for (int i = 0; i < n; i++) {
        char nm[32];
        snprintf(nm, sizeof nm, "chr%d", i + 1);
        img->contigs.push_back(Contig{names ? names[i] : nm, lens[i], 0});
    }
    int rc = image_finalize_layout(img);
    if (rc) { delete img; return rc; }
    for (int ci = 0; ci < n; ci++) {
        const Contig &c = img->contigs[(size_t)ci];
        int64_t b0 = c.goff / kBlockBases, nb = (c.len + 63) / 64;
        for (int64_t b = 0; b < nb; b++) {
            uint64_t lo = splitmix64(seed ^ (uint64_t)(2 * (b0 + b)) * 0xD1342543DE82EF95ull);
            uint64_t hi = splitmix64(seed ^ (uint64_t)(2 * (b0 + b) + 1) * 0xD1342543DE82EF95ull);
            int64_t rem = c.len - b * 64;
            if (rem < 64) { uint64_t m = (1ull << rem) - 1; lo &= m; hi &= m; }
            img->planes[(size_t)(2 * (b0 + b))] = lo;
            img->planes[(size_t)(2 * (b0 + b) + 1)] = hi;
        }
        // N runs: draw (start,len) pairs until the requested fraction is covered, then merge
        int64_t want = (int64_t)(n_frac * (double)c.len);
        std::vector<std::pair<int64_t, int64_t>> rs;
        uint64_t st = splitmix64(seed * 31 + (uint64_t)ci + 7);
        int64_t covered = 0;
        while (covered < want && c.len > run_min) {
            st = splitmix64(st);
            int64_t len = run_min + (int64_t)(st % (uint64_t)(run_max - run_min + 1));
            st = splitmix64(st);
            int64_t s = (int64_t)(st % (uint64_t)c.len);
            int64_t e = std::min(s + len, c.len);
            rs.push_back({s, e});
            covered += e - s;
        }
        std::sort(rs.begin(), rs.end());
        std::vector<std::pair<int64_t, int64_t>> merged;
        for (auto &r : rs) {
            if (!merged.empty() && r.first <= merged.back().second)
                merged.back().second = std::max(merged.back().second, r.second);
            else merged.push_back(r);
        }
        for (auto &r : merged) {
            for (int64_t p = r.first; p < r.second;) {
                int64_t g = c.goff + p, b = g / 64;
                int s = (int)(g & 63);
                int64_t cnt = std::min<int64_t>(64 - s, r.second - p);
                uint64_t m = (cnt == 64 ? ~0ull : ((1ull << cnt) - 1)) << s;
                img->planes[(size_t)(2 * b)] &= ~m;
                img->planes[(size_t)(2 * b + 1)] &= ~m;
                p += cnt;
            }
            img->runs.push_back(Run{(uint32_t)(c.goff + r.first), (uint32_t)(c.goff + r.second), 0, 'N'});
        }
    }
    image_add_pads(img);
    *out = img;
    return GUIDO_OK;
}

int guido_image_free(guido_image *img) {
    GUIDO_API_RANGE();
    if (!img) return GUIDO_OK;
    device_release(img);
    delete img;
    return GUIDO_OK;
}

int guido_image_save(const guido_image *img, const char *path) {
    GUIDO_API_RANGE();
    FILE *f = fopen(path, "wb");
    if (!f) { set_error("cannot write %s", path); return GUIDO_ERR_IO; }
    uint64_t hdr[8] = {kImageMagic, kImageVersion, (uint64_t)img->contigs.size(), (uint64_t)img->runs.size(),
                       (uint64_t)img->n_blocks, (uint64_t)img->padded_bases, (uint64_t)img->total_bases, 0};
    bool ok = fwrite(hdr, sizeof hdr, 1, f) == 1;
    for (auto &c : img->contigs) {
        uint64_t m[3] = {(uint64_t)c.name.size(), (uint64_t)c.len, (uint64_t)c.goff};
        ok = ok && fwrite(m, sizeof m, 1, f) == 1;
        ok = ok && (c.name.empty() || fwrite(c.name.data(), c.name.size(), 1, f) == 1);
    }
    for (auto &r : img->runs) {
        uint32_t m[3] = {r.start, r.end, (uint32_t)r.kind | ((uint32_t)(uint8_t)r.letter << 8)};
        ok = ok && fwrite(m, sizeof m, 1, f) == 1;
    }
    ok = ok && (img->planes.empty() || fwrite(img->planes.data(), 8, img->planes.size(), f) == img->planes.size());
    ok = (fclose(f) == 0) && ok;
    if (!ok) { set_error("short write to %s", path); return GUIDO_ERR_IO; }
    return GUIDO_OK;
}

int guido_image_load(const char *path, guido_image **out) {
    GUIDO_API_RANGE();
    FILE *f = fopen(path, "rb");
    if (!f) { set_error("cannot open %s", path); return GUIDO_ERR_IO; }
    uint64_t hdr[8];
    if (fread(hdr, sizeof hdr, 1, f) != 1 || hdr[0] != kImageMagic || hdr[1] != kImageVersion) {
        fclose(f);
        set_error("%s is not a guido-b200 genome image", path);
        return GUIDO_ERR_IO;
    }
    // the header counts come from the file: nothing is allocated from them before they are checked
    // against the file's size, and no exception may cross the extern "C" boundary
    struct stat sb;
    if (fstat(fileno(f), &sb) != 0) { fclose(f); set_error("cannot stat %s", path); return GUIDO_ERR_IO; }
    const uint64_t fsize = (uint64_t)sb.st_size;
    const bool sane = hdr[2] <= fsize / 24 && hdr[3] <= fsize / 12 && hdr[4] <= fsize / 16 && hdr[5] < (1ull << 32) &&
                      hdr[4] * 64 >= hdr[5] && hdr[6] <= hdr[5];
    if (!sane) { fclose(f); set_error("corrupt image header in %s", path); return GUIDO_ERR_IO; }
    guido_image *img = nullptr;
    bool ok = true;
    try {
        img = new guido_image();
        for (uint64_t i = 0; i < hdr[2] && ok; i++) {
            uint64_t m[3];
            ok = fread(m, sizeof m, 1, f) == 1 && m[0] <= 4096;
            std::string name((size_t)(ok ? m[0] : 0), '\0');
            if (ok && m[0]) ok = fread(&name[0], m[0], 1, f) == 1;
            if (ok) img->contigs.push_back(Contig{name, (int64_t)m[1], (int64_t)m[2]});
        }
        for (uint64_t i = 0; i < hdr[3] && ok; i++) {
            uint32_t m[3];
            ok = fread(m, sizeof m, 1, f) == 1;
            if (ok) img->runs.push_back(Run{m[0], m[1], (uint8_t)(m[2] & 0xff), (char)((m[2] >> 8) & 0xff)});
        }
        img->n_blocks = (int64_t)hdr[4];
        img->padded_bases = (int64_t)hdr[5];
        img->total_bases = (int64_t)hdr[6];
        if (ok) {
            img->planes.resize((size_t)(2 * img->n_blocks));
            ok = img->planes.empty() || fread(img->planes.data(), 8, img->planes.size(), f) == img->planes.size();
        }
    } catch (const std::exception &e) {
        fclose(f);
        delete img;
        set_error("cannot load %s: %s", path, e.what());
        return GUIDO_ERR_NOMEM;
    }
    fclose(f);
    if (!ok) { delete img; set_error("truncated image %s", path); return GUIDO_ERR_IO; }
    *out = img;
    return GUIDO_OK;
}

int guido_image_contig(const guido_image *img, int64_t ix, char *name, int64_t name_cap,
                       int64_t *len, int64_t *goff) {
    GUIDO_API_RANGE();
    if (!img || ix < 0 || ix >= (int64_t)img->contigs.size()) { set_error("contig index out of range"); return GUIDO_ERR_ARG; }
    const Contig &c = img->contigs[(size_t)ix];
    if (name && name_cap > 0) { strncpy(name, c.name.c_str(), (size_t)name_cap - 1); name[name_cap - 1] = 0; }
    if (len) *len = c.len;
    if (goff) *goff = c.goff;
    return GUIDO_OK;
}

int guido_image_fetch(const guido_image *img, int64_t ix, int64_t start0, int64_t n, char *out) {
    GUIDO_API_RANGE();
    if (!img || ix < 0 || ix >= (int64_t)img->contigs.size()) { set_error("contig index out of range"); return GUIDO_ERR_ARG; }
    const Contig &c = img->contigs[(size_t)ix];
    if (start0 < 0 || n < 0 || start0 + n > c.len) { set_error("range outside contig"); return GUIDO_ERR_ARG; }
    static const char L[4] = {'A', 'C', 'G', 'T'};
    for (int64_t i = 0; i < n; i++) {
        int64_t g = c.goff + start0 + i;
        uint64_t lo = img->planes[(size_t)(2 * (g >> 6))], hi = img->planes[(size_t)(2 * (g >> 6) + 1)];
        int s = (int)(g & 63);
        out[i] = L[((lo >> s) & 1) | (((hi >> s) & 1) << 1)];
    }
    // overlay the non-ACGT runs
    uint32_t a = (uint32_t)(c.goff + start0), b = (uint32_t)(c.goff + start0 + n);
    auto it = std::upper_bound(img->runs.begin(), img->runs.end(), a,
                               [](uint32_t v, const Run &r) { return v < r.end; });
    for (; it != img->runs.end() && it->start < b; ++it)
        for (uint32_t p = std::max(a, it->start); p < std::min(b, it->end); p++) out[p - a] = it->letter ? it->letter : 'N';
    return GUIDO_OK;
}

int guido_image_fetch_windows(const guido_image *img, const int64_t *gpos, int64_t n, int32_t width, char *out) {
    GUIDO_API_RANGE();
    if (!img || width < 0 || n < 0) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    static const char L[4] = {'A', 'C', 'G', 'T'};
    for (int64_t w = 0; w < n; w++) {
        char *o = out + w * width;
        int64_t a = gpos[w], b = a + width;
        for (int64_t g = a; g < b; g++) {
            if (g < 0 || g >= img->padded_bases) { o[g - a] = 'N'; continue; }
            uint64_t lo = img->planes[(size_t)(2 * (g >> 6))], hi = img->planes[(size_t)(2 * (g >> 6) + 1)];
            int s = (int)(g & 63);
            o[g - a] = L[((lo >> s) & 1) | (((hi >> s) & 1) << 1)];
        }
        uint32_t ua = (uint32_t)std::max<int64_t>(a, 0), ub = (uint32_t)std::min<int64_t>(b, img->padded_bases);
        auto it = std::upper_bound(img->runs.begin(), img->runs.end(), ua,
                                   [](uint32_t v, const Run &r) { return v < r.end; });
        for (; it != img->runs.end() && it->start < ub; ++it)
            for (uint32_t p = std::max(ua, it->start); p < std::min(ub, it->end); p++) o[(int64_t)p - a] = it->letter ? it->letter : 'N';
    }
    return GUIDO_OK;
}

// the non-ACGT runs (incl. the pads between contigs), sorted by start: [start, end) global, kind 0 = N / pad
int guido_image_get_runs(const guido_image *img, uint32_t *starts, uint32_t *ends, uint8_t *kinds, int64_t cap, int64_t *n_runs) {
    GUIDO_API_RANGE();
    if (!img || !n_runs) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    *n_runs = (int64_t)img->runs.size();
    if ((int64_t)img->runs.size() > cap) return GUIDO_ERR_CAPACITY;
    for (size_t i = 0; i < img->runs.size(); i++) {
        if (starts) starts[i] = img->runs[i].start;
        if (ends) ends[i] = img->runs[i].end;
        if (kinds) kinds[i] = img->runs[i].kind;
    }
    return GUIDO_OK;
}

// 8-byte wire records -> 16-byte records: the window planes come from the host copy of the image,
// cut and oriented exactly as the search kernels do it (offtarget.cu, ot_kernel)
int guido_hits_expand(const guido_image *img, const uint64_t *packed, int64_t n, int32_t pam_len, guido_ot_hit *out) {
    GUIDO_API_RANGE();
    if (!img || n < 0 || pam_len < 1 || pam_len > kMaxPam || (n && (!packed || !out))) { set_error("bad arguments"); return GUIDO_ERR_ARG; }
    const int W = kProto + pam_len;
    const uint32_t wmask = (1u << W) - 1;
    const int64_t n_words = (int64_t)img->planes.size();
    auto rev = [&](uint32_t x) {  // reverse the low W bits
        uint32_t r = 0;
        for (int i = 0; i < W; i++) r |= ((x >> i) & 1u) << (W - 1 - i);
        return r;
    };
    for (int64_t i = 0; i < n; i++) {
        const uint64_t k = packed[i];
        const uint32_t strand = (uint32_t)(k & 1), gpos = (uint32_t)((k >> 1) & 0xffffffffull), guide = (uint32_t)(k >> 33);
        const int64_t b = gpos >> 6;
        const int sh = (int)(gpos & 63);
        if (2 * b + 1 >= n_words) { set_error("hit %lld lies outside the image", (long long)i); return GUIDO_ERR_ARG; }
        uint64_t lo = img->planes[(size_t)(2 * b)] >> sh, hi = img->planes[(size_t)(2 * b + 1)] >> sh;
        if (sh && 2 * b + 3 < n_words) {
            lo |= img->planes[(size_t)(2 * b + 2)] << (64 - sh);
            hi |= img->planes[(size_t)(2 * b + 3)] << (64 - sh);
        }
        uint32_t h = (uint32_t)hi & wmask, l = (uint32_t)lo & wmask;
        if (strand) { h = (rev(h) ^ wmask) | 0x80000000u; l = rev(l) ^ wmask; }
        out[i].guide = guide; out[i].gpos = gpos; out[i].hi = h; out[i].lo = l;
    }
    return GUIDO_OK;
}

int guido_encode_guides(const char *protos, int64_t n, uint32_t *out) {
    GUIDO_API_RANGE();
    uint8_t code[256];  // table lookup: a switch on random letters mispredicts on every second one
    memset(code, 0xff, sizeof code);
    code['A'] = code['a'] = 0; code['C'] = code['c'] = 1; code['G'] = code['g'] = 2; code['T'] = code['t'] = 3;
    for (int64_t g = 0; g < n; g++) {
        uint32_t hi = 0, lo = 0;
        for (int i = 0; i < kProto; i++) {
            const uint32_t c = code[(unsigned char)protos[g * kProto + i]];
            if (c > 3) {
                set_error("guide %lld holds a letter outside A/C/G/T at position %d", (long long)g, i + 1);
                return GUIDO_ERR_ARG;
            }
            lo |= (c & 1) << i;
            hi |= (c >> 1) << i;
        }
        out[2 * g] = hi;
        out[2 * g + 1] = lo;
    }
    return GUIDO_OK;
}

}  // extern "C"
