/*
 * guido_oracle.c -- CPU restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may build or call it.
 * Nothing under guido_b200/ links, imports or executes it.
 *
 * What it restates (citations are into /root/reference):
 *   orc_pam_scan        guido/locus.py:158-183   IUPAC -> look-ahead regex, both strands, overlapping
 *   orc_bowtie_n        bowtie 1.3.1 `-n/-l/-e -a -y` alignment rule as guido invokes it at
 *                       guido/off_targets.py:151-158.  bowtie itself is an un-vendored external
 *                       binary (README.md:17) that is absent from this image, so this function
 *                       restates the bowtie 1 manual's published -n mode semantics:
 *                         * at most `seed_mms` mismatches in the first `seed_len` read bases,
 *                         * sum of (Maq-rounded, capped at 30) qualities over ALL mismatches <= e;
 *                           FASTA reads have quality 'I' (40) -> 30 per mismatch,
 *                         * an N in the read mismatches every reference base,
 *                         * alignments touching an ambiguous reference character, or running off
 *                           the end of a reference sequence, are invalid,
 *                         * both strands, every valid alignment reported (-a, -y).
 *                       PARITY for this one function is pinned only at format level (raw rows in
 *                       test.ipynb:14079-14082, record in docs/source/getting_started.rst:104-112);
 *                       there is no bowtie binary here to run: "parity unpinned" for the bowtie
 *                       step itself, pinned for everything downstream of it.
 *   orc_mmej            guido/mmej.py:8-64        k-mer enumeration + nested-pattern elimination
 *
 * Plain C, no dependencies; OpenMP is used only to let the cpu_baseline use all host cores.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------- helpers */

static inline char up(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }

/* 2-bit code of an upper-case base, -1 for anything that is not A/C/G/T */
static inline int code_of(char c) {
    switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return -1;
    }
}

/* IUPAC letter -> set over {A=1,C=2,G=4,T=8}; guido/locus.py:160-176 (first tuple member).
 * Returns 0 for a letter the reference's dict does not hold (KeyError there). */
static int iupac_set(char c) {
    switch (c) {
    case 'A': return 1;
    case 'C': return 2;
    case 'G': return 4;
    case 'T': return 8;
    case 'R': return 1 | 4;
    case 'Y': return 2 | 8;
    case 'S': return 4 | 2;
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

/* complement of a base set: A<->T, C<->G (second tuple member in locus.py:160-176) */
static int comp_set(int s) {
    return ((s & 1) << 3) | ((s & 2) << 1) | ((s & 4) >> 1) | ((s & 8) >> 3);
}

void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------- PAM scan */

/* guido/locus.py:178-183.  seq is upper-cased on the fly (locus.py:180).  Overlapping matches
 * (look-ahead regex).  fw[i]/rv[i] receive 0-based match starts, ascending.
 * Return 0, or -1 when the PAM holds a letter outside the IUPAC table (KeyError in the
 * reference), or -2 when a capacity is too small (counts are still exact). */
int orc_pam_scan(const char *seq, int64_t n, const char *pam,
                 int32_t *fw, int64_t cap_fw, int64_t *n_fw,
                 int32_t *rv, int64_t cap_rv, int64_t *n_rv) {
    int P = (int)strlen(pam);
    int fset[64], rset[64];
    if (P <= 0 || P > 64) return -1;
    for (int i = 0; i < P; i++) {
        int s = iupac_set(pam[i]);
        if (!s) return -1;
        fset[i] = s;
    }
    for (int i = 0; i < P; i++) rset[i] = comp_set(fset[P - 1 - i]);
    int64_t cf = 0, cr = 0;
    int over = 0;
    for (int64_t p = 0; p + P <= n; p++) {
        int okf = 1, okr = 1;
        for (int i = 0; i < P && (okf || okr); i++) {
            int c = code_of(up(seq[p + i]));
            int bit = (c < 0) ? 0 : (1 << c);
            if (!(fset[i] & bit)) okf = 0;
            if (!(rset[i] & bit)) okr = 0;
        }
        if (okf) { if (cf < cap_fw) fw[cf] = (int32_t)p; else over = 1; cf++; }
        if (okr) { if (cr < cap_rv) rv[cr] = (int32_t)p; else over = 1; cr++; }
    }
    *n_fw = cf;
    *n_rv = cr;
    return over ? -2 : 0;
}

/* --------------------------------------------- bowtie -n mode restatement */

typedef struct {
    int32_t read_ix;
    int32_t contig;
    int64_t offset;     /* 0-based leftmost offset on the forward strand */
    int32_t strand;     /* '+' or '-' as bowtie prints it                 */
    int32_t n_mm;
    uint8_t mm_off[32]; /* offset from the read's 5' end, ascending       */
    char mm_ref[32];    /* reference base, expressed on the forward strand */
    char mm_read[32];   /* read base, expressed on the forward strand      */
} orc_hit;

typedef struct { orc_hit *v; int64_t n, cap; } hitvec;

static void hv_push(hitvec *h, const orc_hit *x) {
    if (h->n == h->cap) {
        h->cap = h->cap ? h->cap * 2 : 1024;
        h->v = (orc_hit *)realloc(h->v, (size_t)h->cap * sizeof(orc_hit));
    }
    h->v[h->n++] = *x;
}

static char comp_base(char c) {
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; }
}

/*
 * reads: n_reads * R characters (upper case, A/C/G/T/N), exactly as guido writes them into its
 * temp FASTA (guido/off_targets.py:132-139).  contigs: raw FASTA sequence text per reference
 * sequence (any case; anything outside ACGT is an ambiguous reference character).
 * Emits every valid alignment; order unspecified.  Returns the number of alignments found
 * (all of them are written when it is <= cap; otherwise the first cap in discovery order).
 */
int64_t orc_bowtie_n(const char *const *contigs, const int64_t *lens, int n_contigs,
                     const char *reads, int n_reads, int R,
                     int seed_mms, int seed_len, int qual_ceiling,
                     orc_hit *out, int64_t cap) {
    if (R <= 0 || R > 32) return -1;
    if (seed_len > R) seed_len = R;
    /* per read and orientation: 2-bit codes, N mask, seed mask -- all in WINDOW coordinates */
    uint64_t *rc_code = (uint64_t *)malloc(sizeof(uint64_t) * 2 * (size_t)n_reads);
    uint64_t *rc_nmask = (uint64_t *)malloc(sizeof(uint64_t) * 2 * (size_t)n_reads);
    uint64_t seedmask[2] = {0, 0};
    for (int o = 0; o < R; o++) {
        if (o < seed_len) {
            seedmask[0] |= 1ull << (2 * o);           /* '+': read offset o sits at window o      */
            seedmask[1] |= 1ull << (2 * (R - 1 - o)); /* '-': read offset o sits at window R-1-o  */
        }
    }
    for (int r = 0; r < n_reads; r++) {
        const char *rd = reads + (size_t)r * R;
        uint64_t c0 = 0, n0 = 0, c1 = 0, n1 = 0;
        for (int o = 0; o < R; o++) {
            int c = code_of(up(rd[o]));
            if (c < 0) { n0 |= 1ull << (2 * o); n1 |= 1ull << (2 * (R - 1 - o)); }
            else {
                c0 |= (uint64_t)c << (2 * o);
                c1 |= (uint64_t)(3 - c) << (2 * (R - 1 - o)); /* reverse complement */
            }
        }
        rc_code[2 * r] = c0; rc_nmask[2 * r] = n0;
        rc_code[2 * r + 1] = c1; rc_nmask[2 * r + 1] = n1;
    }
    const uint64_t LOW = 0x5555555555555555ull;
    const uint64_t wmask = (R == 32) ? ~0ull : ((1ull << (2 * R)) - 1);
    int max_mm = qual_ceiling / 30; /* 30 per mismatch (Maq rounding of 'I'); sum <= e */

    int nthreads = orc_threads();
    hitvec *tv = (hitvec *)calloc((size_t)nthreads, sizeof(hitvec));

    for (int ci = 0; ci < n_contigs; ci++) {
        const char *g = contigs[ci];
        int64_t L = lens[ci];
        if (L < R) continue;
#pragma omp parallel
        {
#ifdef _OPENMP
            int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
            int tid = 0, nt = 1;
#endif
            int64_t nwin = L - R + 1;
            int64_t lo = nwin * tid / nt, hi = nwin * (tid + 1) / nt;
            hitvec *hv = &tv[tid];
            uint64_t w = 0;
            int valid_run = 0; /* number of consecutive unambiguous bases ending at current */
            int64_t pos0 = lo;
            /* prime the rolling window with the first R-1 bases */
            for (int64_t i = pos0; i < pos0 + R - 1 && i < L; i++) {
                int c = code_of(up(g[i]));
                if (c < 0) { valid_run = 0; c = 0; } else valid_run++;
                w = (w >> 2) | ((uint64_t)c << (2 * (R - 1)));
            }
            for (int64_t q = lo; q < hi; q++) {
                int c = code_of(up(g[q + R - 1]));
                if (c < 0) { valid_run = 0; c = 0; } else valid_run++;
                w = ((w >> 2) | ((uint64_t)c << (2 * (R - 1)))) & wmask;
                if (valid_run < R) continue; /* ambiguous reference character in the window */
                for (int r = 0; r < n_reads; r++) {
                    for (int s = 0; s < 2; s++) {
                        uint64_t x = w ^ rc_code[2 * r + s];
                        uint64_t m = ((x | (x >> 1)) & LOW) | rc_nmask[2 * r + s];
                        if (__builtin_popcountll(m) > max_mm) continue;
                        if (__builtin_popcountll(m & seedmask[s]) > seed_mms) continue;
                        orc_hit h;
                        memset(&h, 0, sizeof h);
                        h.read_ix = r; h.contig = ci; h.offset = q; h.strand = s ? '-' : '+';
                        const char *rd = reads + (size_t)r * R;
                        int k = 0;
                        for (int o = 0; o < R; o++) { /* ascending read offset */
                            int wp = s ? (R - 1 - o) : o;
                            if (!((m >> (2 * wp)) & 1)) continue;
                            h.mm_off[k] = (uint8_t)o;
                            h.mm_ref[k] = up(g[q + wp]);
                            char rb = up(rd[o]);
                            h.mm_read[k] = s ? comp_base(rb) : rb;
                            k++;
                        }
                        h.n_mm = k;
                        hv_push(hv, &h);
                    }
                }
            }
        }
    }
    int64_t total = 0;
    for (int t = 0; t < nthreads; t++) {
        for (int64_t i = 0; i < tv[t].n; i++) {
            if (total < cap) out[total] = tv[t].v[i];
            total++;
        }
        free(tv[t].v);
    }
    free(tv); free(rc_code); free(rc_nmask);
    return total;
}

/* pair-compare throughput probe for bench.py's cpu_baseline: same inner loop as above but only
 * counts alignments (no record building), so memory does not bound a long sample. */
int64_t orc_bowtie_n_count(const char *const *contigs, const int64_t *lens, int n_contigs,
                           const char *reads, int n_reads, int R,
                           int seed_mms, int seed_len, int qual_ceiling) {
    return orc_bowtie_n(contigs, lens, n_contigs, reads, n_reads, R, seed_mms, seed_len,
                        qual_ceiling, NULL, 0);
}

/* --------------------------------------------- seed-indexed search (CPU baseline) */

/*
 * The same alignment rule organised the way an index-based aligner runs it, for bench.py's CPU
 * arm: bowtie 1 does not compare a read with every window either -- `bowtie-build` indexes the
 * genome once (guido/genome.py:157-164) and `-n` mode backtracks over the <= n seed mismatches,
 * then extends.  Here: orc_seed_index_build (the bowtie-build counterpart, not timed) tables every
 * window that guido can keep -- N-free, PAM letters matching the fixed letters of the pattern
 * exactly (off_targets.py:169-198), both strands -- under its core k-mer; orc_seed_index_search
 * (timed) enumerates, per read, the cores within core_mm substitutions, and checks the distal
 * bases of the windows filed there against total_mm.  It finds exactly the alignments
 * orc_bowtie_n finds that survive guido's PAM filter (tests/test_oracle_golden.py).
 * Brute force (orc_bowtie_n) stays the parity oracle; this is a throughput baseline.
 */
typedef struct {
    int cl, dl, P;
    uint64_t n;
    uint32_t *off;   /* 4^cl + 1 bucket starts */
    uint32_t *dist;  /* per window: dl distal bases, 2 bits each, guide orientation (base 0 = 5' end) */
    uint64_t *where; /* per window: global position << 1 | strand */
} orc_seed_index;

/* window of W = 20 + P codes in GUIDE orientation (codes[0] = 5' end of the protospacer) -> key / distal */
static inline void split_window(const uint8_t *codes, int cl, int dl, uint32_t *key, uint32_t *dist) {
    uint32_t k = 0, d = 0;
    for (int i = 0; i < dl; i++) d |= (uint32_t)codes[i] << (2 * i);
    for (int i = 0; i < cl; i++) k |= (uint32_t)codes[dl + i] << (2 * i);
    *key = k; *dist = d;
}

/* pam_sets[P]: base sets (A=1,C=2,G=4,T=8) of the PAM in guide orientation; 15 = any */
static int window_at(const char *g, int64_t L, int64_t q, int strand, const int *pam_sets, int P, uint8_t *codes) {
    const int W = 20 + P;
    if (q < 0 || q + W > L) return 0;
    for (int i = 0; i < W; i++) {
        int c = code_of(up(strand ? g[q + W - 1 - i] : g[q + i]));
        if (c < 0) return 0;
        if (strand) c = 3 - c;
        codes[i] = (uint8_t)c;
        if (i >= 20 && !((pam_sets[i - 20] >> c) & 1)) return 0;
    }
    return 1;
}

void orc_seed_index_free(orc_seed_index *ix) {
    if (!ix) return;
    free(ix->off); free(ix->dist); free(ix->where); free(ix);
}

/* goffs[c]: global coordinate of base 0 of contig c (what the caller wants back in `where`) */
orc_seed_index *orc_seed_index_build(const char *const *contigs, const int64_t *lens, const int64_t *goffs,
                                     int n_contigs, const char *pam, int core_len) {
    int P = (int)strlen(pam);
    if (P < 1 || P > 8 || core_len < 1 || core_len > 12) return NULL;
    int pam_sets[8];
    for (int i = 0; i < P; i++) {  /* letters outside A/C/G/T constrain nothing (they are N in the read) */
        char c = up(pam[i]);
        pam_sets[i] = (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? iupac_set(c) : 15;
    }
    orc_seed_index *ix = (orc_seed_index *)calloc(1, sizeof *ix);
    ix->cl = core_len; ix->dl = 20 - core_len; ix->P = P;
    const size_t nb = (size_t)1 << (2 * core_len);
    ix->off = (uint32_t *)calloc(nb + 1, sizeof(uint32_t));
    uint32_t *cnt = (uint32_t *)calloc(nb, sizeof(uint32_t));
    /* pass 1: histogram, pass 2: scatter */
    for (int pass = 0; pass < 2; pass++) {
        for (int ci = 0; ci < n_contigs; ci++) {
            const char *g = contigs[ci];
            const int64_t L = lens[ci];
#pragma omp parallel for schedule(static)
            for (int64_t q = 0; q < L; q++) {
                uint8_t codes[32];
                for (int strand = 0; strand < 2; strand++) {
                    if (!window_at(g, L, q, strand, pam_sets, P, codes)) continue;
                    uint32_t key, dist;
                    split_window(codes, ix->cl, ix->dl, &key, &dist);
                    if (pass == 0) {
#pragma omp atomic
                        cnt[key]++;
                    } else {
                        uint32_t slot;
#pragma omp atomic capture
                        slot = cnt[key]++;
                        ix->dist[slot] = dist;
                        ix->where[slot] = ((uint64_t)(goffs[ci] + q) << 1) | (uint64_t)strand;
                    }
                }
            }
        }
        if (pass == 0) {
            uint64_t run = 0;
            for (size_t k = 0; k < nb; k++) { ix->off[k] = (uint32_t)run; run += cnt[k]; cnt[k] = ix->off[k]; }
            ix->off[nb] = (uint32_t)run;
            ix->n = run;
            if (run >= (1ull << 32)) { free(cnt); orc_seed_index_free(ix); return NULL; }
            ix->dist = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(run + 1));
            ix->where = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(run + 1));
        }
    }
    free(cnt);
    return ix;
}

uint64_t orc_seed_index_size(const orc_seed_index *ix) { return ix ? ix->n : 0; }

/* protos: n_guides x 20 letters (guide orientation).  out (optional, cap records): guide << 33 |
 * position << 1 | strand, unordered.  Returns the number of alignments. */
int64_t orc_seed_index_search(const orc_seed_index *ix, const char *protos, int64_t n_guides, int core_mm,
                              int total_max, uint64_t *out, int64_t cap) {
    const int cl = ix->cl, dl = ix->dl;
    const uint32_t LOW = 0x55555555u;
    int64_t total = 0, cursor = 0;
    if (core_mm > 3) core_mm = 3;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : total)
    for (int64_t gi = 0; gi < n_guides; gi++) {
        uint8_t codes[20];
        int ok = 1;
        for (int i = 0; i < 20; i++) {
            int c = code_of(up(protos[gi * 20 + i]));
            if (c < 0) { ok = 0; break; }
            codes[i] = (uint8_t)c;
        }
        if (!ok) continue;
        uint32_t key = 0, dist = 0;
        for (int i = 0; i < dl; i++) dist |= (uint32_t)codes[i] << (2 * i);
        for (int i = 0; i < cl; i++) key |= (uint32_t)codes[dl + i] << (2 * i);
        int64_t mine = 0;
        /* every core within core_mm substitutions: nm positions, each xor-ed with 1..3 */
        for (int nm = 0; nm <= core_mm; nm++) {
            const int budget = total_max - nm;
            if (budget < 0 || nm > cl) break;
            int pos[3] = {0, 1, 2};
            for (;;) {
                const int n_sub = 1 << (2 * nm);
                for (int sub = 0; sub < n_sub; sub++) {
                    uint32_t k = key;
                    int valid = 1;
                    for (int i = 0; i < nm; i++) {
                        const uint32_t x = (uint32_t)(sub >> (2 * i)) & 3u;
                        if (!x) { valid = 0; break; }
                        k ^= x << (2 * pos[i]);
                    }
                    if (!valid) continue;
                    for (uint32_t s = ix->off[k]; s < ix->off[k + 1]; s++) {
                        const uint32_t x = ix->dist[s] ^ dist;
                        if (__builtin_popcount((x | (x >> 1)) & LOW) > budget) continue;
                        mine++;
                        if (out) {
                            const int64_t at = __atomic_fetch_add(&cursor, 1, __ATOMIC_RELAXED);
                            if (at < cap) out[at] = ((uint64_t)gi << 33) | ix->where[s];
                        }
                    }
                }
                int i = nm - 1;  /* next combination of nm positions out of cl */
                while (i >= 0 && pos[i] == cl - nm + i) i--;
                if (i < 0) break;
                pos[i]++;
                for (int j = i + 1; j < nm; j++) pos[j] = pos[j - 1] + 1;
            }
        }
        total += mine;
    }
    return total;
}

/* ------------------------------------------------------------------ MMEJ */

/* guido/mmej.py:8-64.  seq[0:len] is the window, cut the relative cut position.
 * out receives kept candidates as 4 x uint16 (left_start,left_end,right_start,right_end) in the
 * reference's order.  *n_cand = number of candidates before elimination.
 * Returns the number kept (written only up to cap), or -1 when there is no candidate at all
 * (the reference returns None, mmej.py:119-120). */
typedef struct { int ls, le, rs, re; } cand_t;

static int occurs(const char *hay, int n, const char *needle, int k) {
    for (int j = 0; j + k <= n; j++)
        if (memcmp(hay + j, needle, (size_t)k) == 0) return 1;
    return 0;
}

/* greedy left-to-right non-overlapping occurrences, like re.finditer (mmej.py:38-39) */
static int find_all(const char *hay, int n, const char *needle, int k, int *starts) {
    int c = 0, j = 0;
    while (j + k <= n) {
        if (memcmp(hay + j, needle, (size_t)k) == 0) { starts[c++] = j; j += k; }
        else j++;
    }
    return c;
}

int64_t orc_mmej(const char *seq, int len, int cut, uint16_t *out, int64_t cap, int64_t *n_cand) {
    if (cut < 0) cut = 0; /* callers never pass this; Python slicing would differ */
    if (cut > len) cut = len;
    const char *left = seq;
    int nl = cut;
    const char *right = seq + cut;
    int nr = len - cut;
    cand_t *cands = NULL;
    int64_t nc = 0, ccap = 0;
    int *acc_i = (int *)malloc(sizeof(int) * (size_t)(nl + 1));
    int *lpos = (int *)malloc(sizeof(int) * (size_t)(nl + 1));
    int *rpos = (int *)malloc(sizeof(int) * (size_t)(nr + 1));
    for (int k = nl - 1; k >= 2; k--) {           /* mmej.py:24 */
        int n_acc = 0;                            /* accepted k-mers of this length */
        for (int i = 0; i + k <= nl; i++) {       /* mmej.py:27 */
            const char *kmer = left + i;
            if (!occurs(right, nr, kmer, k)) continue;           /* `kmer in right_seq` */
            int seen = 0;                                        /* `kmer not in kmers` */
            for (int a = 0; a < n_acc && !seen; a++)
                if (memcmp(left + acc_i[a], kmer, (size_t)k) == 0) seen = 1;
            if (seen) continue;
            int cl = find_all(left, nl, kmer, k, lpos);
            int cr = find_all(right, nr, kmer, k, rpos);
            for (int a = 0; a < cl; a++)
                for (int b = 0; b < cr; b++) {    /* itertools.product, left-major */
                    if (nc == ccap) {
                        ccap = ccap ? ccap * 2 : 256;
                        cands = (cand_t *)realloc(cands, (size_t)ccap * sizeof(cand_t));
                    }
                    cands[nc].ls = lpos[a]; cands[nc].le = lpos[a] + k;
                    cands[nc].rs = rpos[b]; cands[nc].re = rpos[b] + k;
                    nc++;
                }
            acc_i[n_acc++] = i;
        }
    }
    free(acc_i); free(lpos); free(rpos);
    *n_cand = nc;
    if (nc == 0) { free(cands); return -1; }
    /* mmej.py:54-64: j is a sub-pattern iff some earlier i contains it on both sides */
    int64_t kept = 0;
    for (int64_t j = 0; j < nc; j++) {
        int sub = 0;
        for (int64_t i = 0; i < j && !sub; i++)
            if (cands[i].ls <= cands[j].ls && cands[i].le >= cands[j].le &&
                cands[i].rs <= cands[j].rs && cands[i].re >= cands[j].re) sub = 1;
        if (sub) continue;
        if (kept < cap) {
            out[4 * kept + 0] = (uint16_t)cands[j].ls; out[4 * kept + 1] = (uint16_t)cands[j].le;
            out[4 * kept + 2] = (uint16_t)cands[j].rs; out[4 * kept + 3] = (uint16_t)cands[j].re;
        }
        kept++;
    }
    free(cands);
    return kept;
}
