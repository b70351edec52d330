/*
 * miekki_oracle.c -- CPU restatement of Miekki's sketch-and-query hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under miekki_b200/ (the product) may
 * include, link or call this file.  It is used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs as the checker (or the timed CPU baseline), never as the product path.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors
 * (SURVEY.md section 4), so this restatement is pinned against outputs of the
 * unmodified reference binary compiled by oracle/Makefile into
 * oracle/_ref/Miekki: fingerprint rows, sketch_size, genome_size, Bloom bytes
 * (through the -d dump), hit lines (-a) and exact-mode lines (-e).  The
 * vectors and the script that made them are in tests/golden/.
 *
 * Every function cites the reference lines it restates (paths are relative to
 * the reference tree).  Written from the behaviour described in SURVEY.md
 * Appendix A; plain C99, scalar, single threaded.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MKO_EMPTY_FP 255u               /* Miekki.cpp:29 maximal_minimizer */
#define MKO_EMPTY_ANC (~(uint64_t)0)    /* Miekki.cpp:30 maximal_hash      */

/* ---- character codes ---------------------------------------------------- */

/* utils.cpp:31-49 nuc2int: C,G,T -> 1,2,3 ; every other byte -> 0 */
static inline uint64_t fwd_code(unsigned char c) {
    return c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 0u;
}

/* utils.cpp:107-125 nuc2intrc: A,C,G -> 3,2,1 ; every other byte -> 0 */
static inline uint64_t rc_code(unsigned char c) {
    return c == 'A' ? 3u : c == 'C' ? 2u : c == 'G' ? 1u : 0u;
}

/* utils.cpp:252-272 str2numstrand: case-insensitive 2-bit packing, returns 0
 * as soon as a byte outside ACGTacgt is met. */
uint64_t mko_str2numstrand(const char *s, uint64_t n) {
    uint64_t res = 0;
    for (uint64_t i = 0; i < n; ++i) {
        res <<= 2;
        switch (s[i]) {
            case 'A': case 'a': break;
            case 'C': case 'c': res += 1; break;
            case 'G': case 'g': res += 2; break;
            case 'T': case 't': res += 3; break;
            default: return 0;
        }
    }
    return res;
}

/* ---- hashes ------------------------------------------------------------- */

/* utils.cpp:179-184 */
uint64_t mko_revhash64(uint64_t x) {
    x = ((x >> 32) ^ x) * 0xD6E8FEB86659FD93ull;
    x = ((x >> 32) ^ x) * 0xD6E8FEB86659FD93ull;
    x = ((x >> 32) ^ x);
    return x;
}

/* utils.cpp:188-193 */
uint64_t mko_unrevhash64(uint64_t x) {
    x = ((x >> 32) ^ x) * 0xCFEE444D8B59A89Bull;
    x = ((x >> 32) ^ x) * 0xCFEE444D8B59A89Bull;
    x = ((x >> 32) ^ x);
    return x;
}

/* utils.cpp:197-199: k*69 is a 32-bit product, widened for the 64-bit one */
uint64_t mko_universal_hash(uint64_t x, uint32_t i) {
    uint32_t m = i * 69u;
    return mko_unrevhash64(x) + (((uint64_t)m * mko_revhash64(x)) % 1024u);
}

/* ---- fingerprint -------------------------------------------------------- */

/* Miekki.cpp:91-113 mantis + utils.cpp:84-91 asm_log2 (bsr = index of the top
 * set bit).  nbm = number_bit_minimizer, nbmant = number_bit_mantis. */
uint8_t mko_mantis(uint64_t n, uint32_t h, uint32_t nbm, uint32_t nbmant) {
    if (n == 0) return (uint8_t)MKO_EMPTY_FP;
    int64_t prefix = 63 - __builtin_clzll(n);
    int64_t e = prefix - 32 + (int64_t)h;
    if (e < 0) e = 0;
    int offset = (int)(prefix - (int64_t)(nbm - nbmant));
    if (offset < 0) offset = 0;
    uint64_t suffix = n - ((uint64_t)1 << prefix);
    suffix >>= offset;
    uint64_t res = suffix + ((uint64_t)e << (nbm - nbmant));
    return (uint8_t)res;
}

/* ---- rolling k-mers ----------------------------------------------------- */

/* Miekki.cpp:66-76 rcb: reverse complement of a packed k-mer (k digits) */
static uint64_t rcb(uint64_t m, uint32_t k) {
    uint64_t res = 0, offset = (uint64_t)1 << (2 * k - 2);
    for (uint32_t i = 0; i < k; ++i) {
        res += (3 - (m % 4)) * offset;
        m >>= 2;
        offset >>= 2;
    }
    return res;
}

/* Miekki.cpp:150-197 minhash_sketch_partition.
 * fp[B] <- 255, anc[B] <- ~0, then the loop "i + k < n" (n-k iterations: the
 * last k-mer is never visited, Miekki.cpp:162).  Returns active buckets. */
uint32_t mko_sketch(const char *s, uint64_t n, uint32_t k, uint32_t h,
                    uint32_t nbm, uint32_t nbmant, uint8_t *fp, uint64_t *anc) {
    const uint64_t B = (uint64_t)1 << h;
    const uint64_t part = (uint64_t)1 << (64 - h);           /* :151 */
    const uint64_t kmask = ((uint64_t)1 << (2 * k)) - 1;     /* Miekki.h:76-77 */
    memset(fp, 0xFF, B);
    for (uint64_t b = 0; b < B; ++b) anc[b] = MKO_EMPTY_ANC;
    uint32_t active = 0;
    uint64_t plen = n < (uint64_t)(k - 1) ? n : (uint64_t)(k - 1);  /* substr(0,k-1) :158 */
    uint64_t S = mko_str2numstrand(s, plen);
    uint64_t RC = rcb(S, k);                                 /* :160 */
    for (uint64_t i = 0; i + k < n; ++i) {                   /* :162 */
        unsigned char c = (unsigned char)s[i + k - 1];
        S = ((S << 2) + fwd_code(c)) & kmask;                /* :51-55 */
        RC = (RC >> 2) + (rc_code(c) << (2 * k - 2));        /* :59-62 */
        uint64_t x = mko_revhash64(S < RC ? S : RC);         /* :167-168 */
        uint64_t bucket = x / part;                          /* :169 */
        uint8_t v = mko_mantis(x % part, h, nbm, nbmant);    /* :170-171 */
        if (v < fp[bucket]) {                                /* :172 strict */
            if (fp[bucket] == MKO_EMPTY_FP) ++active;
            fp[bucket] = v;
            anc[bucket] = x;
        }
    }
    return active;
}

/* ---- Bloom filter ------------------------------------------------------- */

/* Miekki.cpp:135-146 check_bloom: byte-granular ("cell && mask[hit]" is a
 * logical AND, so only cell != 0 matters). */
int mko_bloom_check(const uint8_t *table, uint32_t b, uint64_t x) {
    for (uint32_t i = 0; i < 5; ++i) {
        uint64_t slot = mko_universal_hash(x, i) >> b;
        if (table[slot / 8] == 0) return 0;
    }
    return 1;
}

/* Miekki.cpp:121-131 insert_bloom */
void mko_bloom_insert(uint8_t *table, uint32_t b, uint64_t x) {
    for (uint32_t i = 0; i < 5; ++i) {
        uint64_t slot = mko_universal_hash(x, i) >> b;
        if (table[slot / 8] == 0) table[slot / 8] += (uint8_t)(1u << (slot % 8));
    }
}

/* Highest Bloom byte index + 1 that a k-mer of size k can touch (the window the
 * product keeps on device): canonical k-mer < 4^k, plus < 1024 from the i term. */
uint64_t mko_bloom_window(uint32_t k, uint32_t b) {
    uint64_t top = (k >= 32) ? ~(uint64_t)0 : (((uint64_t)1 << (2 * k)) + 1023);
    return ((top >> b) / 8) + 1;
}

/* ---- index -------------------------------------------------------------- */

typedef struct {
    uint32_t k, h, nbm, nbmant, b;
    uint64_t B;
    uint32_t n, cap;          /* genomes inserted / capacity                    */
    uint8_t *rows;            /* bucket-major: rows[bucket * cap + genome]      */
    uint32_t *sketch_size;    /* Miekki.h:59 */
    uint64_t *genome_size;    /* Miekki.h:60 */
    uint8_t *bloom;           /* 2^b / 8 bytes, Miekki.h:56,85-89               */
    uint64_t bloom_bytes;
    uint8_t *tmp_fp;
    uint64_t *tmp_anc;
} mko_index;

mko_index *mko_index_new(uint32_t k, uint32_t h, uint32_t nbm, uint32_t nbmant,
                         uint32_t b, uint32_t cap, uint64_t bloom_bytes) {
    mko_index *ix = (mko_index *)calloc(1, sizeof(mko_index));
    ix->k = k; ix->h = h; ix->nbm = nbm; ix->nbmant = nbmant; ix->b = b;
    ix->B = (uint64_t)1 << h;
    ix->cap = cap ? cap : 1;
    ix->rows = (uint8_t *)malloc(ix->B * ix->cap);
    memset(ix->rows, 0xFF, ix->B * ix->cap);
    ix->sketch_size = (uint32_t *)calloc(ix->cap, sizeof(uint32_t));
    ix->genome_size = (uint64_t *)calloc(ix->cap, sizeof(uint64_t));
    /* callers may pass a shortened table (>= mko_bloom_window) to save RAM */
    ix->bloom_bytes = bloom_bytes ? bloom_bytes : (((uint64_t)1 << b) / 8);
    ix->bloom = (uint8_t *)calloc(ix->bloom_bytes, 1);
    ix->tmp_fp = (uint8_t *)malloc(ix->B);
    ix->tmp_anc = (uint64_t *)malloc(ix->B * sizeof(uint64_t));
    return ix;
}

void mko_index_free(mko_index *ix) {
    if (!ix) return;
    free(ix->rows); free(ix->sketch_size); free(ix->genome_size);
    free(ix->bloom); free(ix->tmp_fp); free(ix->tmp_anc); free(ix);
}

uint32_t mko_index_n(const mko_index *ix) { return ix->n; }
uint32_t mko_index_cap(const mko_index *ix) { return ix->cap; }
uint8_t *mko_index_rows(mko_index *ix) { return ix->rows; }
uint32_t *mko_index_sketch_size(mko_index *ix) { return ix->sketch_size; }
uint64_t *mko_index_genome_size(mko_index *ix) { return ix->genome_size; }
uint8_t *mko_index_bloom(mko_index *ix) { return ix->bloom; }
uint64_t mko_index_bloom_bytes(const mko_index *ix) { return ix->bloom_bytes; }
void mko_index_set_n(mko_index *ix, uint32_t n) { ix->n = n; }

/* Miekki.cpp:287-311: body of insert_sequences for one genome (ids are
 * insertion order).  Returns the genome id or -1 when full. */
int64_t mko_index_insert(mko_index *ix, const char *s, uint64_t n) {
    if (ix->n >= ix->cap) return -1;
    const uint32_t g = ix->n;
    mko_sketch(s, n, ix->k, ix->h, ix->nbm, ix->nbmant, ix->tmp_fp, ix->tmp_anc);
    double approx = 0;              /* :288 */
    uint32_t active = 0;            /* :289 uint32_t */
    const uint32_t sh = ix->nbm - ix->nbmant;
    for (uint64_t i = 0; i < ix->B; ++i) {
        ix->rows[i * ix->cap + g] = ix->tmp_fp[i];          /* :291 add_index */
        if (ix->tmp_fp[i] != MKO_EMPTY_FP) {
            /* :293  1/pow(2, fp >> 3): an exact power of two */
            approx += 1.0 / (double)((uint64_t)1 << (ix->tmp_fp[i] >> sh));
            ++active;
            if (!mko_bloom_check(ix->bloom, ix->b, ix->tmp_anc[i]))   /* :295 */
                mko_bloom_insert(ix->bloom, ix->b, ix->tmp_anc[i]);   /* :298 */
        }
    }
    ix->sketch_size[g] = active;                            /* :305 */
    /* :306  active*active is a uint32_t product and wraps (quirk G2) */
    uint32_t sq = active * active;
    approx = (0.72134 * (double)sq) / approx;
    if (approx > (double)n) ix->genome_size[g] = n;         /* :307-308 */
    else ix->genome_size[g] = (uint64_t)approx;             /* :310 */
    ix->n = g + 1;
    return g;
}

/* Miekki.cpp:214-224 + :344-372: Bloom-masked sketch of one read, then
 * count[g] = #{bucket : fp != 255 and row[bucket][g] == fp}.
 * If masked_fp != NULL it receives the B masked fingerprints.
 * Returns A(q) = number of surviving buckets. */
uint32_t mko_query_counts(mko_index *ix, const char *s, uint64_t n,
                          uint32_t *counts, uint8_t *masked_fp) {
    mko_sketch(s, n, ix->k, ix->h, ix->nbm, ix->nbmant, ix->tmp_fp, ix->tmp_anc);
    uint32_t surviving = 0;
    memset(counts, 0, (size_t)ix->n * sizeof(uint32_t));
    for (uint64_t i = 0; i < ix->B; ++i) {
        /* :216-219: check_bloom runs on every bucket; an empty bucket holds
         * anc = ~0 and stays 255 whatever the answer. */
        uint8_t f = ix->tmp_fp[i];
        if (f != MKO_EMPTY_FP && !mko_bloom_check(ix->bloom, ix->b, ix->tmp_anc[i]))
            f = MKO_EMPTY_FP;
        if (masked_fp) masked_fp[i] = f;
        if (f == MKO_EMPTY_FP) continue;                    /* :359-361 */
        ++surviving;
        const uint8_t *row = ix->rows + i * ix->cap;
        for (uint32_t g = 0; g < ix->n; ++g)                /* :363-366 */
            counts[g] += (row[g] == f);
    }
    return surviving;
}

/* ---- threshold + top-k (libstdc++ heap semantics) ----------------------- */

typedef struct {
    uint32_t genome, matches;       /* Miekki.h:27-31 similarity_score */
    double jaccard, intersection;
} mko_hit;

/* comp(a,b) := a.intersection > b.intersection   (Miekki.cpp:377) */
#define COMP(a, b) ((a).intersection > (b).intersection)

/* bits/stl_heap.h:128-149 __push_heap */
static void heap_push_up(mko_hit *a, long hole, long top, mko_hit v) {
    long parent = (hole - 1) / 2;
    while (hole > top && COMP(a[parent], v)) {
        a[hole] = a[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a[hole] = v;
}

/* bits/stl_heap.h:224-250 __adjust_heap */
static void heap_adjust(mko_hit *a, long hole, long len, mko_hit v) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (COMP(a[child], a[child - 1])) child--;
        a[hole] = a[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a[hole] = a[child - 1];
        hole = child - 1;
    }
    heap_push_up(a, hole, top, v);
}

/* bits/stl_heap.h:254-266 __pop_heap on a[0..n): moves the top to a[n-1] */
static void heap_pop(mko_hit *a, long n) {
    mko_hit v = a[n - 1];
    a[n - 1] = a[0];
    heap_adjust(a, 0, n - 1, v);
}

/* Miekki.cpp:376-397 filter_results.  hits must hold nresults entries.
 * first_id lets a shard score columns [first_id, first_id+N) of a wider index;
 * heap/len_io, when non-NULL, carry the heap state in and out (no final sort)
 * so shards can be chained in ascending id order. */
uint32_t mko_filter_chain(const uint32_t *counts, uint32_t N, uint32_t first_id,
                          const uint32_t *sketch_size, const uint64_t *genome_size,
                          uint32_t nresults, uint32_t min_score, double min_intersection,
                          mko_hit *heap, uint32_t len, int finalize) {
    for (uint32_t g = 0; g < N; ++g) {
        uint32_t score = counts[g];
        if (score < min_score) continue;                              /* :381 */
        double jaccard = (double)score / (double)sketch_size[g];      /* :382 */
        double inter = jaccard * (double)genome_size[g];              /* :383 */
        if (inter < min_intersection) continue;                       /* :384 */
        if (len >= nresults) {                                        /* :386 */
            if (heap[0].intersection > inter) continue;               /* :387 */
            heap_pop(heap, len);                                      /* :388 */
            --len;                                                    /* :389 */
        }
        mko_hit v = { first_id + g, score, jaccard, inter };          /* :392 */
        heap[len++] = v;
        heap_push_up(heap, len - 1, 0, v);                            /* :393 */
    }
    if (finalize)                                                     /* :396 sort_heap */
        for (long n = len; n > 1; --n) heap_pop(heap, n);
    return len;
}

uint32_t mko_filter(const uint32_t *counts, uint32_t N,
                    const uint32_t *sketch_size, const uint64_t *genome_size,
                    uint32_t nresults, uint32_t min_score, double min_intersection,
                    mko_hit *hits) {
    return mko_filter_chain(counts, N, 0, sketch_size, genome_size, nresults,
                            min_score, min_intersection, hits, 0, 1);
}

/* ---- exact mode --------------------------------------------------------- */

/* utils.cpp:203-214 revCompChar */
static inline char rev_comp_char(char c) {
    switch (c) {
        case 'C': case 'c': return 'G';
        case 'G': case 'g': return 'C';
        case 'T': case 't': return 'A';
    }
    return 'T';
}

/* utils.cpp:276-278 str2num: min(str2numstrand(w), str2numstrand(revComp(w))) */
uint64_t mko_str2num(const char *w, uint32_t k) {
    char rc[64];
    for (uint32_t i = 0; i < k; ++i) rc[k - 1 - i] = rev_comp_char(w[i]);  /* :218-224 */
    uint64_t a = mko_str2numstrand(w, k), b = mko_str2numstrand(rc, k);
    return a < b ? a : b;
}

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

/* sort + unique in place; returns the distinct count */
static uint64_t sort_unique(uint64_t *v, uint64_t n) {
    if (n == 0) return 0;
    qsort(v, n, sizeof(uint64_t), cmp_u64);
    uint64_t m = 1;
    for (uint64_t i = 1; i < n; ++i)
        if (v[i] != v[m - 1]) v[m++] = v[i];
    return m;
}

/* Miekki.cpp:803-822: the set B of canonical k-mers of a genome file given as
 * its records (already joined per record the way the reference's loop leaves
 * them, see oracle.py:records_like_reference).  Returns a malloc'd sorted
 * distinct array and its size. */
uint64_t *mko_exact_genome_set(const char *const *recs, const uint64_t *lens,
                               uint32_t nrec, uint32_t k, uint64_t *out_n) {
    uint64_t total = 0;
    for (uint32_t r = 0; r < nrec; ++r)
        if (lens[r] >= k) total += lens[r] - k + 1;
    uint64_t *v = (uint64_t *)malloc((total ? total : 1) * sizeof(uint64_t));
    uint64_t m = 0;
    for (uint32_t r = 0; r < nrec; ++r) {
        if (lens[r] < k) continue;                                    /* :806,:817 */
        for (uint64_t i = 0; i + k - 1 < lens[r]; ++i)                /* :807 */
            v[m++] = mko_str2num(recs[r] + i, k);
    }
    *out_n = sort_unique(v, m);
    return v;
}

void mko_free(void *p) { free(p); }

/* Miekki.cpp:826-842: for one read, nb_inter = |A n B| and
 * nb_union = |B| + |A \ B| over the distinct canonical k-mers A of the read. */
void mko_exact_read(const uint64_t *setB, uint64_t nB, const char *s, uint64_t n,
                    uint32_t k, uint64_t *nb_inter, uint64_t *nb_union) {
    uint64_t inter = 0, uni = nB;
    if (n >= k) {
        uint64_t cnt = n - k + 1;
        uint64_t *a = (uint64_t *)malloc(cnt * sizeof(uint64_t));
        for (uint64_t i = 0; i < cnt; ++i) a[i] = mko_str2num(s + i, k);
        uint64_t m = sort_unique(a, cnt);
        for (uint64_t i = 0; i < m; ++i) {
            uint64_t key = a[i];
            if (nB && bsearch(&key, setB, nB, sizeof(uint64_t), cmp_u64)) ++inter;
            else ++uni;
        }
        free(a);
    }
    *nb_inter = inter;
    *nb_union = uni;
}
