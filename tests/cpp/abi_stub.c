/*
 * abi_stub.c -- TEST INFRASTRUCTURE, not product code.
 *
 * A CPU stand-in for libmiekki_b200.so: the entry points of include/miekki_b200.h that the
 * `miekki` command line calls, implemented over the oracle (oracle/miekki_oracle.c).  The CPU
 * suite links the unmodified CLI source against it (tests/test_cli_stub_cpu.py) so that the
 * host logic -- option handling, FASTA rules, batching, heap chaining over shards, hit-line /
 * exact-line formatting, the streamed gz dump and its loader -- is checked against the
 * reference binary's golden outputs without a GPU.  Nothing here is shipped or measured; the
 * product library has no CPU path (mk_create fails without a B200).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/miekki_b200.h"

/* oracle API (oracle/miekki_oracle.c) */
typedef struct mko_index mko_index;
typedef struct { uint32_t genome, matches; double jaccard, intersection; } mko_hit;
mko_index *mko_index_new(uint32_t k, uint32_t h, uint32_t nbm, uint32_t nbmant, uint32_t b, uint32_t cap,
                         uint64_t bloom_bytes);
void mko_index_free(mko_index *ix);
uint32_t mko_index_n(const mko_index *ix);
uint32_t mko_index_cap(const mko_index *ix);
uint8_t *mko_index_rows(mko_index *ix);
uint32_t *mko_index_sketch_size(mko_index *ix);
uint64_t *mko_index_genome_size(mko_index *ix);
uint8_t *mko_index_bloom(mko_index *ix);
uint64_t mko_index_bloom_bytes(const mko_index *ix);
void mko_index_set_n(mko_index *ix, uint32_t n);
int64_t mko_index_insert(mko_index *ix, const char *s, uint64_t n);
uint32_t mko_query_counts(mko_index *ix, const char *s, uint64_t n, uint32_t *counts, uint8_t *masked_fp);
uint32_t mko_filter_chain(const uint32_t *counts, uint32_t N, uint32_t first_id, const uint32_t *sketch_size,
                          const uint64_t *genome_size, uint32_t nresults, uint32_t min_score,
                          double min_intersection, mko_hit *heap, uint32_t len, int finalize);
uint64_t mko_bloom_window(uint32_t k, uint32_t b);
uint64_t *mko_exact_genome_set(const char *const *recs, const uint64_t *lens, uint32_t nrec, uint32_t k,
                               uint64_t *out_n);
void mko_exact_read(const uint64_t *setB, uint64_t nB, const char *s, uint64_t n, uint32_t k,
                    uint64_t *nb_inter, uint64_t *nb_union);
void mko_free(void *p);

struct mk_ctx {
    uint32_t k, h, nbm, nbmant, b, threshold, first_id;
    uint64_t B, window;
    mko_index *ix;
    uint32_t importing;
    int import_open;
    uint32_t *counts;          /* mk_scan: n_reads x N */
    uint32_t scan_reads;
    char err[256];
};
struct mk_batch {
    uint32_t n;
    char **seq;
    uint64_t *len;
};

static char g_err[256];
/* the oracle index keeps per-index scratch: calls that sketch are serialised (the product
 * library serialises per ctx as well) */
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;

static int fail(mk_ctx *c, int code, const char *msg) {
    snprintf(c ? c->err : g_err, 256, "%s", msg);
    return code;
}

/* the oracle index has a fixed capacity: move to a wider one when needed */
static void ensure_cap(mk_ctx *c, uint32_t need) {
    const uint32_t cap = mko_index_cap(c->ix), n = mko_index_n(c->ix);
    if (need <= cap) return;
    uint32_t ncap = cap * 2 > need ? cap * 2 : need;
    mko_index *nx = mko_index_new(c->k, c->h, c->nbm, c->nbmant, c->b, ncap, c->window);
    for (uint64_t r = 0; r < c->B; ++r)
        memcpy(mko_index_rows(nx) + r * ncap, mko_index_rows(c->ix) + r * cap, n);
    memcpy(mko_index_sketch_size(nx), mko_index_sketch_size(c->ix), (size_t)n * 4);
    memcpy(mko_index_genome_size(nx), mko_index_genome_size(c->ix), (size_t)n * 8);
    memcpy(mko_index_bloom(nx), mko_index_bloom(c->ix), c->window);
    mko_index_set_n(nx, n);
    mko_index_free(c->ix);
    c->ix = nx;
}

int mk_abi_version(void) { return MK_ABI_VERSION; }
const char *mk_last_error(const mk_ctx *c) { return c ? c->err : g_err; }

int mk_create(uint32_t k, uint32_t h, uint32_t bits_per_min, uint32_t bits_mantis, uint32_t bloom_log2,
              uint32_t threshold, int device, mk_ctx **out) {
    if (!out) return fail(NULL, MK_ERR_ARG, "out is NULL");
    *out = NULL;
    if (bits_per_min != 8 || bits_mantis != 5) return fail(NULL, MK_ERR_UNSUPPORTED, "not implemented");
    if (k < 2 || k > 31 || h < 1 || h > 24 || bloom_log2 < 32 || bloom_log2 > 40 || device < 0)
        return fail(NULL, MK_ERR_ARG, "bad parameter");
    mk_ctx *c = (mk_ctx *)calloc(1, sizeof(mk_ctx));
    c->k = k; c->h = h; c->nbm = bits_per_min; c->nbmant = bits_mantis; c->b = bloom_log2;
    c->threshold = threshold;
    c->B = (uint64_t)1 << h;
    c->window = (mko_bloom_window(k, bloom_log2) + 15) / 16 * 16;
    c->ix = mko_index_new(k, h, bits_per_min, bits_mantis, bloom_log2, 16, c->window);
    *out = c;
    return MK_OK;
}

void mk_destroy(mk_ctx *c) {
    if (!c) return;
    mko_index_free(c->ix);
    free(c->counts);
    free(c);
}

int mk_set_shard(mk_ctx *c, uint32_t first_id) {
    c->first_id = first_id;
    return MK_OK;
}

int mk_index_size(const mk_ctx *c, uint32_t *n) {
    *n = mko_index_n(c->ix);
    return MK_OK;
}

int mk_index_reserve(mk_ctx *c, uint32_t n_genomes) {
    pthread_mutex_lock(&g_mu);
    ensure_cap(c, n_genomes);
    pthread_mutex_unlock(&g_mu);
    return MK_OK;
}

int mk_index_add(mk_ctx *c, const char *const *seqs, const uint64_t *lens, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i)
        if (lens[i] < c->k) return fail(c, MK_ERR_ARG, "mk_index_add: sequence shorter than k");
    pthread_mutex_lock(&g_mu);
    ensure_cap(c, mko_index_n(c->ix) + n);
    for (uint32_t i = 0; i < n; ++i) mko_index_insert(c->ix, seqs[i], lens[i]);
    pthread_mutex_unlock(&g_mu);
    return MK_OK;
}

int mk_index_export_rows(mk_ctx *c, uint64_t row0, uint64_t nrows, uint8_t *dst, uint64_t dst_stride) {
    const uint32_t n = mko_index_n(c->ix), cap = mko_index_cap(c->ix);
    if (row0 + nrows > c->B || (nrows && dst_stride < n)) return fail(c, MK_ERR_ARG, "mk_index_export_rows: bad range");
    for (uint64_t r = 0; r < nrows; ++r) memcpy(dst + r * dst_stride, mko_index_rows(c->ix) + (row0 + r) * cap, n);
    return MK_OK;
}

int mk_index_export(mk_ctx *c, uint8_t *rows, uint64_t *genome_size, uint8_t *bloom, uint64_t bloom_bytes,
                    uint32_t *sketch_size) {
    const uint32_t n = mko_index_n(c->ix);
    if (rows) mk_index_export_rows(c, 0, c->B, rows, n);
    if (genome_size) memcpy(genome_size, mko_index_genome_size(c->ix), (size_t)n * 8);
    if (sketch_size) memcpy(sketch_size, mko_index_sketch_size(c->ix), (size_t)n * 4);
    if (bloom) {
        const uint64_t m = bloom_bytes < c->window ? bloom_bytes : c->window;
        memcpy(bloom, mko_index_bloom(c->ix), m);
        if (bloom_bytes > m) memset(bloom + m, 0, bloom_bytes - m);
    }
    return MK_OK;
}

int mk_index_import_begin(mk_ctx *c, uint32_t n) {
    mko_index_free(c->ix);
    c->ix = mko_index_new(c->k, c->h, c->nbm, c->nbmant, c->b, n ? n : 1, c->window);
    c->importing = n;
    c->import_open = 1;
    return MK_OK;
}

int mk_index_import_rows(mk_ctx *c, uint64_t row0, uint64_t nrows, const uint8_t *src, uint64_t src_stride) {
    if (!c->import_open) return fail(c, MK_ERR_STATE, "mk_index_import_rows: call mk_index_import_begin first");
    const uint32_t cap = mko_index_cap(c->ix);
    if (row0 + nrows > c->B) return fail(c, MK_ERR_ARG, "mk_index_import_rows: bad range");
    for (uint64_t r = 0; r < nrows; ++r)
        memcpy(mko_index_rows(c->ix) + (row0 + r) * cap, src + r * src_stride, c->importing);
    return MK_OK;
}

int mk_index_import_end(mk_ctx *c, const uint64_t *genome_size, const uint8_t *bloom, uint64_t bloom_bytes,
                        const uint32_t *sketch_size) {
    if (!c->import_open) return fail(c, MK_ERR_STATE, "mk_index_import_end: call mk_index_import_begin first");
    const uint32_t n = c->importing;
    memcpy(mko_index_genome_size(c->ix), genome_size, (size_t)n * 8);
    memcpy(mko_index_sketch_size(c->ix), sketch_size, (size_t)n * 4);
    memset(mko_index_bloom(c->ix), 0, c->window);
    if (bloom && bloom_bytes) memcpy(mko_index_bloom(c->ix), bloom, bloom_bytes < c->window ? bloom_bytes : c->window);
    mko_index_set_n(c->ix, n);
    c->import_open = 0;
    return MK_OK;
}

/* Miekki.cpp:901-906 (columns appended) + the Bloom fold of SURVEY.md 8e (ours wins where set) */
int mk_index_merge(mk_ctx *c, mk_ctx *other) {
    if (!c || !other || c == other) return fail(c, MK_ERR_ARG, "mk_index_merge: bad argument");
    if (c->k != other->k || c->h != other->h || c->b != other->b)
        return fail(c, MK_ERR_ARG, "mk_index_merge: the two indexes were built with different -k / -h / -f / -b");
    pthread_mutex_lock(&g_mu);
    const uint32_t n = mko_index_n(c->ix), m = mko_index_n(other->ix), ocap = mko_index_cap(other->ix);
    ensure_cap(c, n + m);
    const uint32_t cap = mko_index_cap(c->ix);
    for (uint64_t r = 0; r < c->B; ++r)
        memcpy(mko_index_rows(c->ix) + r * cap + n, mko_index_rows(other->ix) + r * ocap, m);
    memcpy(mko_index_sketch_size(c->ix) + n, mko_index_sketch_size(other->ix), (size_t)m * 4);
    memcpy(mko_index_genome_size(c->ix) + n, mko_index_genome_size(other->ix), (size_t)m * 8);
    uint8_t *mine = mko_index_bloom(c->ix);
    const uint8_t *theirs = mko_index_bloom(other->ix);
    for (uint64_t i = 0; i < c->window; ++i)
        if (!mine[i]) mine[i] = theirs[i];
    mko_index_set_n(c->ix, n + m);
    pthread_mutex_unlock(&g_mu);
    return MK_OK;
}

uint64_t mk_bloom_window(const mk_ctx *c) { return c->window; }
int mk_bloom_get(mk_ctx *c, uint8_t *dst, uint64_t n) {
    memcpy(dst, mko_index_bloom(c->ix), n);
    return MK_OK;
}
int mk_bloom_set(mk_ctx *c, const uint8_t *src, uint64_t n) {
    memcpy(mko_index_bloom(c->ix), src, n);
    return MK_OK;
}

int mk_batch_upload(mk_ctx *c, const char *const *seqs, const uint64_t *lens, uint32_t n, mk_batch **out) {
    (void)c;
    mk_batch *b = (mk_batch *)calloc(1, sizeof(mk_batch));
    b->n = n;
    b->seq = (char **)calloc(n ? n : 1, sizeof(char *));
    b->len = (uint64_t *)calloc(n ? n : 1, sizeof(uint64_t));
    for (uint32_t i = 0; i < n; ++i) {
        b->seq[i] = (char *)malloc(lens[i] + 1);
        memcpy(b->seq[i], seqs[i], lens[i]);
        b->len[i] = lens[i];
    }
    *out = b;
    return MK_OK;
}

void mk_batch_free(mk_ctx *c, mk_batch *b) {
    (void)c;
    if (!b) return;
    for (uint32_t i = 0; i < b->n; ++i) free(b->seq[i]);
    free(b->seq);
    free(b->len);
    free(b);
}

static void counts_of(mk_ctx *c, const char *s, uint64_t len, uint32_t *counts) {
    const uint32_t N = mko_index_n(c->ix);
    if (len < c->k) memset(counts, 0, (size_t)N * 4);          /* the reference's callers skip these */
    else mko_query_counts(c->ix, s, len, counts, NULL);
}

int mk_scan(mk_ctx *c, const mk_batch *reads) {
    const uint32_t N = mko_index_n(c->ix);
    free(c->counts);
    c->counts = (uint32_t *)calloc((size_t)reads->n * (N ? N : 1) + 1, 4);
    c->scan_reads = reads->n;
    pthread_mutex_lock(&g_mu);
    for (uint32_t i = 0; i < reads->n; ++i) counts_of(c, reads->seq[i], reads->len[i], c->counts + (size_t)i * N);
    pthread_mutex_unlock(&g_mu);
    return MK_OK;
}

int mk_topk(mk_ctx *c, uint32_t nresults, uint32_t min_score, double min_intersection, mk_hit *heap_io,
            uint32_t *len_io, int chain_in, int finalize) {
    const uint32_t N = mko_index_n(c->ix);
    for (uint32_t i = 0; i < c->scan_reads; ++i)
        len_io[i] = mko_filter_chain(c->counts + (size_t)i * N, N, c->first_id, mko_index_sketch_size(c->ix),
                                     mko_index_genome_size(c->ix), nresults, min_score, min_intersection,
                                     (mko_hit *)(heap_io + (size_t)i * nresults), chain_in ? len_io[i] : 0, finalize);
    return MK_OK;
}

int mk_query(mk_ctx *c, const char *const *seqs, const uint64_t *lens, uint32_t n, uint32_t nresults,
             uint32_t min_score, double min_intersection, mk_hit *hits, uint32_t *nhits) {
    const uint32_t N = mko_index_n(c->ix);
    uint32_t *counts = (uint32_t *)calloc(N ? N : 1, 4);
    pthread_mutex_lock(&g_mu);
    for (uint32_t i = 0; i < n; ++i) {
        counts_of(c, seqs[i], lens[i], counts);
        nhits[i] = mko_filter_chain(counts, N, c->first_id, mko_index_sketch_size(c->ix),
                                    mko_index_genome_size(c->ix), nresults, min_score, min_intersection,
                                    (mko_hit *)(hits + (size_t)i * nresults), 0, 1);
    }
    pthread_mutex_unlock(&g_mu);
    free(counts);
    return MK_OK;
}

int mk_exact(mk_ctx *c, const char *const *records, const uint64_t *rec_lens, uint32_t n_records,
             const char *const *reads, const uint64_t *read_lens, uint32_t n_reads, uint64_t *nb_inter,
             uint64_t *nb_union, uint64_t *genome_distinct) {
    uint64_t nB = 0;
    uint64_t *setB = mko_exact_genome_set(records, rec_lens, n_records, c->k, &nB);
    if (genome_distinct) *genome_distinct = nB;
    for (uint32_t i = 0; i < n_reads; ++i)
        mko_exact_read(setB, nB, reads[i], read_lens[i], c->k, nb_inter + i, nb_union + i);
    mko_free(setB);
    return MK_OK;
}

int mk_exact_many(mk_ctx *c, uint32_t n_genomes, const char *const *records, const uint64_t *rec_lens,
                  const uint32_t *rec_count, const char *const *reads, const uint64_t *read_lens,
                  const uint32_t *read_count, uint64_t *nb_inter, uint64_t *nb_union, uint64_t *genome_distinct) {
    uint64_t r0 = 0, q0 = 0;
    for (uint32_t g = 0; g < n_genomes; ++g) {
        uint64_t nB = 0;
        int rc = mk_exact(c, records + r0, rec_lens + r0, rec_count[g], reads + q0, read_lens + q0, read_count[g],
                          nb_inter + q0, nb_union + q0, &nB);
        if (rc != MK_OK) return rc;
        if (genome_distinct) genome_distinct[g] = nB;
        r0 += rec_count[g];
        q0 += read_count[g];
    }
    return MK_OK;
}
