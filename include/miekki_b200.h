/*
 * miekki_b200.h -- C ABI of the B200-native sketch-and-query path of Miekki.
 *
 * The reference (Malfoy/Miekki) has no FFI: its boundary is `class Miekki`
 * (Miekki.h:33-137) as driven by main.cpp:184-235.  Each entry point below
 * names the reference method it replaces.  Host code (the `miekki` CLI in
 * miekki_b200/cli/, or any binding -- see INTEGRATION.md) keeps FASTA/gz
 * parsing and text formatting and calls these with plain pointers and sizes.
 *
 * Conventions
 *  - every call returns 0 on success, <0 on error (mk_last_error() has the text);
 *  - the caller owns every host buffer; nothing returned points into the library
 *    except the error string;
 *  - one mk_ctx drives one GPU (one process per GPU in multi-GPU runs: each rank
 *    owns a contiguous ascending range of genome ids, see mk_set_shard);
 *  - calls on one ctx are serialised internally, so host parser threads may call
 *    concurrently; work is enqueued on the ctx's CUDA stream and every call
 *    returns after its results are in the caller's buffers;
 *  - there is no CPU fallback: without a CUDA device mk_create fails.
 */
#ifndef MIEKKI_B200_H
#define MIEKKI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MK_ABI_VERSION 1

enum {
    MK_OK = 0,
    MK_ERR_ARG = -1,       /* bad argument (see mk_last_error)                     */
    MK_ERR_CUDA = -2,      /* CUDA runtime error                                    */
    MK_ERR_NOMEM = -3,     /* host or device allocation failed                      */
    MK_ERR_UNSUPPORTED = -4, /* e.g. -f other than 3: reference prints "not implemented" (Miekki.cpp:236) */
    MK_ERR_STATE = -5      /* call not valid in this state                          */
};

typedef struct mk_ctx mk_ctx;

/* == similarity_score, Miekki.h:27-31 (24 bytes, same field order) */
typedef struct mk_hit {
    uint32_t genome;        /* insertion id (list order), not a file name */
    uint32_t matches;       /* shared fingerprints                        */
    double jaccard;         /* matches / sketch_size[genome]              */
    double intersection;    /* jaccard * genome_size[genome]              */
} mk_hit;

typedef struct mk_batch mk_batch;   /* sequences resident in HBM */

/* ---- lifetime -------------------------------------------------------------- */

/* Miekki::Miekki(k, h, bits_per_min, bits_mantis, 0, out, b, threshold, t), Miekki.h:66.
 * bits_per_min must be 8 (the -f 3 default; quirk G12) and bits_mantis 5.
 * k in [2,31] (Miekki.h:76-77), h in [1,24], bloom_log2 in [32,40] (smaller values
 * index out of bounds in the reference, Miekki.cpp:124).  device = CUDA ordinal. */
int mk_create(uint32_t k, uint32_t h, uint32_t bits_per_min, uint32_t bits_mantis,
              uint32_t bloom_log2, uint32_t threshold, int device, mk_ctx **out);
void mk_destroy(mk_ctx *ctx);
const char *mk_last_error(const mk_ctx *ctx);   /* ctx may be NULL: last mk_create error */
int mk_abi_version(void);

/* Run all later work of this ctx on an existing CUDA stream (a cudaStream_t passed as
 * void*, e.g. torch.cuda.current_stream().cuda_stream).  NULL restores the ctx's own. */
int mk_set_stream(mk_ctx *ctx, void *cuda_stream);

/* Genome-sharded runs: this ctx holds global ids
 * [first_id, first_id + mk_index_size).  Hits carry global ids. */
int mk_set_shard(mk_ctx *ctx, uint32_t first_id);

/* The scan is a persistent kernel that fills every SM.  A sharded run passes the heap from GPU
 * to GPU (NCCL send/recv) while the next batch is being scanned: leave n SMs out of the scan's
 * grid so that those kernels are scheduled at once instead of after the scan (default 0). */
int mk_set_scan_spare_sms(mk_ctx *ctx, int n);

/* parameters as stored in a dump header (Miekki.cpp:651-661) */
int mk_get_params(const mk_ctx *ctx, uint32_t *k, uint32_t *h, uint32_t *bits_per_min,
                  uint32_t *bits_mantis, uint32_t *bloom_log2, uint32_t *threshold);

/* ---- sequences in HBM ------------------------------------------------------ */

/* Copies n ASCII sequences (any bytes; no terminator needed) to the device. */
int mk_batch_upload(mk_ctx *ctx, const char *const *seqs, const uint64_t *lens, uint32_t n,
                    mk_batch **out);
/* Same, from one contiguous host buffer (ideally pinned): sequence i is
 * data[offsets[i] .. offsets[i] + lens[i]).  When every offset is the previous one plus the
 * length rounded up to 16 (offsets[0] == 0) the whole buffer goes over in a single copy. */
int mk_batch_upload_flat(mk_ctx *ctx, const char *data, const uint64_t *offsets,
                         const uint64_t *lens, uint32_t n, mk_batch **out);
/* Bench/test tooling: n synthetic genomes of `len` bases generated on the device with the
 * counter-based generator of miekki_b200/synth.py:cb_bases (ids first_g .. first_g+n-1). */
int mk_batch_synth(mk_ctx *ctx, uint64_t seed, uint32_t first_g, uint32_t n, uint64_t len,
                   mk_batch **out);
int mk_batch_download(mk_ctx *ctx, const mk_batch *b, uint32_t i, char *dst, uint64_t cap);
uint32_t mk_batch_size(const mk_batch *b);
uint64_t mk_batch_bases(const mk_batch *b);
void mk_batch_free(mk_ctx *ctx, mk_batch *b);

/* ---- build ----------------------------------------------------------------- */

/* Pre-size the bucket-major matrix for n_genomes columns (optional; it grows). */
int mk_index_reserve(mk_ctx *ctx, uint32_t n_genomes);

/* Miekki::insert_sequences, Miekki.cpp:277-314.  Appends n genomes; ids are assigned in
 * call order (== `-t 1` list order).  Sequences shorter than k are an error, as the
 * reference's callers filter them (Miekki.cpp:569). */
int mk_index_add(mk_ctx *ctx, const char *const *seqs, const uint64_t *lens, uint32_t n);
int mk_index_add_batch(mk_ctx *ctx, const mk_batch *genomes);   /* same, input already in HBM */

int mk_index_size(const mk_ctx *ctx, uint32_t *n_genomes);
/* sketch_size (Miekki.h:59) and genome_size (Miekki.h:60) of ids [first, first+n) (local ids) */
int mk_index_stats(mk_ctx *ctx, uint32_t first, uint32_t n, uint32_t *sketch_size,
                   uint64_t *genome_size);

/* Payload of Miekki::dump_disk, Miekki.cpp:665-676: rows is 2^h x N bucket-major (row b =
 * N bytes in id order, no padding); bloom receives bloom_bytes (= 2^b / 8) bytes.  Any
 * pointer may be NULL to skip that part. */
int mk_index_export(mk_ctx *ctx, uint8_t *rows, uint64_t *genome_size, uint8_t *bloom,
                    uint64_t bloom_bytes, uint32_t *sketch_size);
/* Loader ctor Miekki::Miekki(const string&), Miekki.cpp:682-719 (after the header was
 * parsed and passed to mk_create).  Replaces the index content.  rows_stride is the
 * distance in bytes between consecutive bucket rows in `rows` (>= n). */
int mk_index_import(mk_ctx *ctx, uint32_t n, const uint8_t *rows, uint64_t rows_stride,
                    const uint64_t *genome_size, const uint8_t *bloom, uint64_t bloom_bytes,
                    const uint32_t *sketch_size);

/* Streaming forms of the two calls above, for indexes that should not sit in host memory as
 * one block (a -h 20 index of 10,000 genomes is 10.5 GB): the dump is written / read a slab
 * of bucket rows at a time, in the order of the file (rows, then statistics and Bloom table).
 * export_rows: rows [row0, row0 + nrows) to dst, row r at dst + (r - row0) * dst_stride
 * (dst_stride >= N lets several shards fill their columns of one host slab).
 * import: begin(n) empties the index and sizes it for n genomes; rows arrive in any number
 * of import_rows calls; end() installs statistics and Bloom bytes and publishes the index. */
int mk_index_export_rows(mk_ctx *ctx, uint64_t row0, uint64_t nrows, uint8_t *dst, uint64_t dst_stride);
int mk_index_import_begin(mk_ctx *ctx, uint32_t n);
int mk_index_import_rows(mk_ctx *ctx, uint64_t row0, uint64_t nrows, const uint8_t *src,
                         uint64_t src_stride);
int mk_index_import_end(mk_ctx *ctx, const uint64_t *genome_size, const uint8_t *bloom,
                        uint64_t bloom_bytes, const uint32_t *sketch_size);

/* Miekki::merge_indexes, Miekki.cpp:901-910: appends the genomes of `other` after ours (ids
 * mk_index_size(ctx) ...), with their statistics, and folds the two Bloom tables "ours wins where
 * set" -- the reference leaves that as a TODO (:907).  The result equals one in-order build of
 * both genome lists.  Both indexes must share k, h, fingerprint width and bloom_log2; they may
 * live on different GPUs.  `other` is left unchanged and still owned by the caller (the reference
 * deletes it, :908). */
int mk_index_merge(mk_ctx *ctx, mk_ctx *other);

/* Bloom bytes only (multi-GPU merge): the window the device keeps is the
 * first mk_bloom_window() bytes of the 2^b/8-byte table; bytes past it are never touched
 * for this k.  merge: dst byte = (dst != 0) ? dst : src  ("lowest rank wins"). */
uint64_t mk_bloom_window(const mk_ctx *ctx);
/* Bytes [mk_bloom_reach(k, b), window) can never be probed by a k-mer of size k: an upper bound
 * of the canonical k-mer (see api.cu).  Pure arithmetic, needs no device. */
uint64_t mk_bloom_reach(uint32_t k, uint32_t bloom_log2);
int mk_bloom_get(mk_ctx *ctx, uint8_t *dst, uint64_t n);
int mk_bloom_merge(mk_ctx *ctx, const uint8_t *src, uint64_t n);
int mk_bloom_set(mk_ctx *ctx, const uint8_t *src, uint64_t n);   /* replace the first n bytes */
/* dst/src of the three calls above may be host or device pointers (unified addressing).  A device
 * buffer must be complete when the call is made: the copy runs on the context's stream (see
 * mk_set_stream), which does not wait for work queued on other streams. */

/* ---- query ----------------------------------------------------------------- */

/* Miekki::query_sequences + filter_results, Miekki.cpp:344-372 + :376-422.
 * hits: n * nresults entries, read i at hits[i*nresults ..], nhits[i] valid ones, sorted
 * like std::sort_heap leaves them (descending intersection; libstdc++ tie order).
 * Reads shorter than k get nhits = 0 (the reference's callers skip them, :465).
 * `-a` uses (10, 10, 0.5*threshold), `-e` (5, 10, threshold). */
int mk_query(mk_ctx *ctx, const char *const *seqs, const uint64_t *lens, uint32_t n,
             uint32_t nresults, uint32_t min_score, double min_intersection,
             mk_hit *hits, uint32_t *nhits);
/* Same with reads already in HBM; hits/nhits are host buffers (NULL: leave on device). */
int mk_query_batch(mk_ctx *ctx, const mk_batch *reads, uint32_t nresults, uint32_t min_score,
                   double min_intersection, mk_hit *hits, uint32_t *nhits);

/* Sharded top-k (SURVEY.md 7 hard part 1): ranks are chained in ascending genome-id order.
 * heap_io: n * nresults mk_hit, len_io: n lengths -- the bounded heap of Miekki.cpp:386-393
 * as the previous shard left it (len 0 for the first shard).  finalize != 0 on the last
 * shard applies std::sort_heap (:396). */
int mk_query_chain(mk_ctx *ctx, const mk_batch *reads, uint32_t nresults, uint32_t min_score,
                   double min_intersection, mk_hit *heap_io, uint32_t *len_io, int finalize);

/* The two halves of mk_query_chain, so that shards scan concurrently and only the cheap
 * top-k step is chained: mk_scan sketches the reads and scans this shard, keeping the
 * count matrix (n x N u32) in HBM; mk_topk then applies Miekki.cpp:376-397 to it.
 * chain_in != 0: heap_io/len_io hold the previous shard's state.  heap_io/len_io may be
 * host or device pointers.  mk_scan fails with MK_ERR_ARG when n x N counts exceed 16 GiB
 * (split the reads). */
int mk_scan(mk_ctx *ctx, const mk_batch *reads);
int mk_topk(mk_ctx *ctx, uint32_t nresults, uint32_t min_score, double min_intersection,
            mk_hit *heap_io, uint32_t *len_io, int chain_in, int finalize);
/* Pipelined form of the pair above.  mk_scan_async only enqueues the sketch + scan on the ctx
 * stream and returns at once; counts land in one of two tiles, *slot says which.
 * mk_topk_slot waits (on the device) for that scan, runs the heap step on a second stream and
 * returns when heap_io/len_io are written, without waiting for later scans -- so batch i+1
 * scans while batch i's heap travels through the shards.  At most two batches in flight: call
 * mk_topk_slot for a slot before its tile is scanned into again.  `reads` must stay alive
 * until the scan has run (e.g. until the matching mk_topk_slot returned). */
int mk_scan_async(mk_ctx *ctx, const mk_batch *reads, int *slot);
/* Optional, ahead of mk_scan_async: enqueues the read sketch of the batch the NEXT mk_scan_async
 * will be given (which then only enqueues the scan), so that a caller who has to wait for
 * something else before it may scan -- the previous batch's heap from another shard -- does not
 * leave the sketch for afterwards.  Returns at once.  The index must not change in between; any
 * other batch passed to mk_scan_async is simply sketched there as usual. */
int mk_sketch_async(mk_ctx *ctx, const mk_batch *reads);
int mk_topk_slot(mk_ctx *ctx, int slot, uint32_t nresults, uint32_t min_score,
                 double min_intersection, mk_hit *heap_io, uint32_t *len_io, int chain_in,
                 int finalize);
/* mk_topk_slot for reads [first_read, first_read + n_reads) of the slot's batch only (clipped to
 * the batch); heap_io / len_io address read first_read.  A chain over shards passes a batch on in
 * tiles of reads this way: shard r steps tile t while shard r + 1 steps tile t - 1, and the last
 * batch of a run drains in one heap step plus one tile per further shard instead of one whole
 * heap step per shard. */
int mk_topk_slot_range(mk_ctx *ctx, int slot, uint32_t first_read, uint32_t n_reads,
                       uint32_t nresults, uint32_t min_score, double min_intersection,
                       mk_hit *heap_io, uint32_t *len_io, int chain_in, int finalize);

/* Test hook: raw shared-fingerprint counts, the matrix Miekki::query_sequences returns
 * (Miekki.cpp:352): counts[i * N + g].  surviving (may be NULL) receives A(q), the number of
 * query buckets that are non-empty and pass the Bloom check. */
int mk_query_counts(mk_ctx *ctx, const char *const *seqs, const uint64_t *lens, uint32_t n,
                    uint32_t *counts, uint32_t *surviving);

/* Test hook: Miekki::minhash_sketch_partition, Miekki.cpp:150-197.  fp[2^h], anc[2^h]. */
int mk_sketch(mk_ctx *ctx, const char *seq, uint64_t len, uint8_t *fp, uint64_t *anc,
              uint32_t *active);

/* ---- exact mode (-e) ------------------------------------------------------- */

/* Miekki::ground_truth_batch, Miekki.cpp:792-859, for one genome file: `records` are the
 * strings whose k-mers form set B (the host applies the record rules of :803-822), reads
 * the candidates grouped on that genome.  Per read: nb_inter = |A n B|,
 * nb_union = |B| + |A \ B| over distinct canonical k-mers (str2num, utils.cpp:276). */
int mk_exact(mk_ctx *ctx, const char *const *records, const uint64_t *rec_lens, uint32_t n_records,
             const char *const *reads, const uint64_t *read_lens, uint32_t n_reads,
             uint64_t *nb_inter, uint64_t *nb_union, uint64_t *genome_distinct);
/* The same for n_genomes genome files in one call (query_file_exact flushes its candidates genome
 * by genome, Miekki.cpp:745-754): genome g owns the next rec_count[g] entries of records / rec_lens
 * and the next read_count[g] entries of reads / read_lens; nb_inter / nb_union are per read in that
 * order, genome_distinct (may be NULL) per genome.  One upload, no host round trip between genomes. */
int mk_exact_many(mk_ctx *ctx, uint32_t n_genomes, const char *const *records, const uint64_t *rec_lens,
                  const uint32_t *rec_count, const char *const *reads, const uint64_t *read_lens,
                  const uint32_t *read_count, uint64_t *nb_inter, uint64_t *nb_union, uint64_t *genome_distinct);
/* Same as mk_exact with the records and the reads already in HBM (mk_batch_upload / mk_batch_synth). */
int mk_exact_batch(mk_ctx *ctx, const mk_batch *records, const mk_batch *reads,
                   uint64_t *nb_inter, uint64_t *nb_union, uint64_t *genome_distinct);

/* ---- measurement ----------------------------------------------------------- */

/* Device-side timings (CUDA events on the ctx stream) and work counters accumulated since
 * the last mk_stats_reset.  Times in milliseconds. */
typedef struct mk_stats {
    double sketch_ms;          /* encode + sketch + finalize kernels (build)             */
    double read_sketch_ms;     /* query-side sketch + Bloom mask                         */
    double scan_ms;            /* fingerprint scan kernel                                */
    double topk_ms;            /* threshold + bounded-heap kernel                        */
    double exact_ms;
    uint64_t scan_launches;
    uint64_t scan_row_bytes;   /* algorithmic bytes: sum_q A(q) * N_shard                */
    uint64_t scan_rows;        /* sum_q A(q)                                             */
    uint64_t bases_sketched;   /* genome bases through the build kernels                 */
    uint64_t bases_queried;    /* read bases through the query kernels                   */
    uint64_t kernel_launches;  /* every kernel this library launched                     */
    uint64_t h2d_bytes, d2h_bytes;
} mk_stats;
int mk_stats_get(mk_ctx *ctx, mk_stats *out);
int mk_stats_reset(mk_ctx *ctx);
int mk_sync(mk_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* MIEKKI_B200_H */
