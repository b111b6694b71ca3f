// join_blocks.cuh -- interface of the bit-sliced block join (join_blocks.cu) used by offtarget.cu.
#pragma once
#include "device.cuh"

namespace guido {

constexpr int kSiteBlockSlots = 64;  // windows per bit-sliced block
constexpr int kSiteBlockWords = 80;  // 32-bit words per block: 10 distal bases x 4 letters x 64 windows, one bit each

struct BjArgs {
    const uint32_t *blocks;   // site table: kSiteBlockWords words per block of 64 slots
    const uint32_t *blk_off;  // n_buckets + 1: first SLOT of each bucket (multiple of 64) | pad count in the low 6 bits
    const uint4 *slots;       // per slot {global position, hi plane | strand << 31, lo plane, core key}; pads: x = ~0
    const uint2 *ent;         // seed index entries {distal letters (2 bits each) | bias << 20, guide}, bucket order;
                              // inside a bucket the entries of the largest bias class come first
    const uint32_t *ent_off;  // n_buckets + 1: first entry of each bucket (even) | 1 if an unused pad entry follows the bucket
    uint32_t key_lo, key_hi;  // buckets joined by this launch
    uint32_t group, n_groups; // buckets per ticket, tickets (set by launch_join_blocks)
    uint32_t *ticket;         // zero at launch
    guido_ot_hit *hits;
    unsigned long long cap;
    unsigned long long *count;
    unsigned long long *site_count;  // [1] += entries x slots compared (may be null)
};

int launch_join_blocks(DeviceState *ds, BjArgs a, cudaStream_t stream);
int launch_site_blocks(DeviceState *ds, const uint4 *d_slots, uint32_t n_blocks, int dl, uint32_t *d_blocks,
                       cudaStream_t stream);

}  // namespace guido
