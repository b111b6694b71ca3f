// internal.h -- shared declarations of libguido_b200 (not part of the C-ABI).
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/guido_b200.h"

namespace guido {

constexpr int kBlockBases = 64;     // one {lo,hi} u64 pair per 64 bases
constexpr int kProto = 20;          // protospacer length the reference assumes (locus.py:316)
constexpr int kMaxPam = 8;          // 20 + P must fit the 28 payload bits of guido_ot_hit
constexpr uint32_t kImageMagic = 0x42324447u;  // "GD2B"
constexpr uint32_t kImageVersion = 1;

struct Contig {
    std::string name;
    int64_t len;
    int64_t goff;  // global coordinate of base 0
};

// maximal stretch of one non-ACGT letter (after upper-casing), or a pad between contigs
struct Run {
    uint32_t start;  // global, inclusive
    uint32_t end;    // global, exclusive
    uint8_t kind;    // 0: 'N' or pad (dropped on both strands), 1: any other letter
    char letter;     // 0 for pads
};

// PAM compiled to base sets (bit0 A, bit1 C, bit2 G, bit3 T), guido/locus.py:160-179
struct PamSpec {
    int32_t len;
    uint8_t fw[kMaxPam];   // set at PAM position i, forward strand
    uint8_t rv[kMaxPam];   // set at window position i when the PAM is on the reverse strand
    int32_t n_ambiguous;   // letters that rev_comp() turns into 'N' in the bowtie read
    uint32_t fixed_mask;   // bit i set: PAM letter i is one of A/C/G/T
    uint8_t code[kMaxPam]; // 2-bit code of fixed letters
};

struct DeviceState;  // defined in device.cu

}  // namespace guido

struct guido_image {
    std::vector<guido::Contig> contigs;
    std::vector<guido::Run> runs;     // sorted by start, disjoint
    std::vector<uint64_t> planes;     // 2 * n_blocks words: [2b] lo plane, [2b+1] hi plane
    int64_t n_blocks = 0;
    int64_t padded_bases = 0;
    int64_t total_bases = 0;
    guido::DeviceState *dev = nullptr;
};

namespace guido {

// NVTX range around a C-ABI entry point (visible in nsys / ncu timelines; a no-op without a tool
// attached: nvtx3 is header-only and resolves its injection library lazily)
struct ApiRange {
    explicit ApiRange(const char *name);
    ~ApiRange();
};
#define GUIDO_API_RANGE() ::guido::ApiRange guido_api_range_(__func__)

void set_error(const char *fmt, ...);
int parse_pam(const char *pam, PamSpec *spec);  // GUIDO_ERR_PAM on unknown letter
int image_finalize_layout(guido_image *img);     // assigns goff / padded_bases / n_blocks
void image_pack_contig(guido_image *img, size_t contig_ix, const char *seq, int64_t len);
void image_add_pads(guido_image *img);
int64_t image_find_contig(const guido_image *img, int64_t gpos);
void device_release(guido_image *img);

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

}  // namespace guido
