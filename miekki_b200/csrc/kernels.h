// kernels.h -- host-side launchers of the sm_100a kernels (one .cu per kernel family).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mk {

struct SketchParams {
    int k, h;
    uint32_t bloom_log2;
    uint64_t bloom_window;      // bytes of the Bloom table kept on the device
};

// Opt a kernel in to `bytes` of dynamic shared memory on the CURRENT device.  The attribute is
// per device and per function; the size already granted is remembered per (device, function)
// under a mutex, so launches from several host threads / on several GPUs are safe.
bool smem_optin(const void* func, size_t bytes);

// ---- sketch.cu ------------------------------------------------------------------
constexpr int KEY_TAG_BITS = 24;    // tagged sketch keys, see launch_sketch_dense
// Dense path (genomes, long reads): sequences -> 2-bit planes -> per-bucket min key.
//   chars     : concatenated ASCII, sequence s at chars + coff[s] (16-byte aligned)
//   woff[s]   : first plane word of sequence s (each sequence owns ceil(len/16)+2 words)
//   planes    : interleaved, planeR == planeF + 1 and word w of a plane sits at index 2 w
//   keys      : n_seq x 2^h u64, pre-set to ~0; afterwards fp << 56 | first position << ks | tag
//   ks        : 0 (plain keys) or KEY_TAG_BITS (every sequence shorter than 2^32: the low bits
//               carry the top bits of the canonical k-mer, see resolve_kernel)
void launch_encode_planes(const uint8_t* chars, const uint64_t* coff, const uint64_t* len,
                          const uint64_t* woff, uint32_t n_seq, uint64_t max_len, int k,
                          uint32_t* planeF, uint32_t* planeR, cudaStream_t st);
// Planes of sequences packed on the host (pack.h).  packed[woff[s] + w] = forward digits of word w
// of sequence s (whole words of upper-case ACGT outside the prefix); exc = the other words as raw
// bytes (seq relative to the batch, seq0 = first sequence of this view); pvalid[s] = the k-1
// prefix of sequence s is all ACGTacgt.
struct PackExcDev {
    uint4 bytes;
    uint64_t word;
    uint32_t seq, pad;
};
void launch_expand_planes(const uint32_t* packed, const PackExcDev* exc, uint32_t n_exc, uint32_t seq0,
                          const uint8_t* pvalid, const uint64_t* len, const uint64_t* woff, uint32_t n_seq,
                          uint64_t max_len, int k, uint32_t* planes, cudaStream_t st);
void launch_sketch_dense(const uint32_t* planeF, const uint32_t* planeR, const uint64_t* len,
                         const uint64_t* woff, uint32_t n_seq, uint64_t max_len, int k, int h,
                         unsigned long long* keys, int ks, cudaStream_t st);
// keys -> fp[s][b] (u8) and anc written over keys[s][b]; per-sequence active count and
// S = sum 2^(31 - (fp >> 3)).  If owner != nullptr, also registers Bloom candidates
// (pass A of the order-exact insert) with key (seq_in_batch, bucket, probe); then anc is only
// written for buckets that may still change the Bloom table (pair_full from launch_bloom_pages;
// nullptr: every bucket is looked up).
void launch_resolve(unsigned long long* keys_anc, const uint32_t* planeF, const uint32_t* planeR,
                    const uint64_t* woff, uint32_t n_seq, SketchParams p, uint8_t* fp,
                    uint32_t* active, unsigned long long* ssum, const uint8_t* bloom,
                    uint32_t* owner, const uint32_t* pair_full, uint32_t pair_words, int ks,
                    uint32_t* claims, uint32_t* n_claims, cudaStream_t st);
// The same for a build with tagged keys (ks != 0) and a page bitmap, as two kernels: a streaming
// pass over all keys (fp, statistics, list of buckets on unsaturated pages) and a look-up pass
// over that list, which leaves the claimants in it (other entries become 0xFFFFFFFF).
bool resolve_tagged_ok(int h, uint32_t pair_words);
void launch_resolve_tagged(unsigned long long* keys_anc, const uint32_t* planeF, const uint64_t* woff,
                           uint32_t n_seq, SketchParams p, uint8_t* fp, uint32_t* active, unsigned long long* ssum,
                           const uint8_t* bloom, uint32_t* owner, const uint32_t* pair_full, uint32_t pair_words,
                           int ks, uint32_t* list, uint32_t* n_list, cudaStream_t st);
// Bloom pages without a zero byte -> full8[n_pages], pair_full[ceil(n_pages / 32)]
uint32_t bloom_page_count(uint64_t window);
// reach: bytes [reach, window) can never be probed and do not count
uint32_t launch_bloom_pages(const uint8_t* bloom, uint64_t window, uint64_t reach, uint8_t* full8,
                            uint32_t* pair_full, cudaStream_t st);
// Miekki.cpp:303-311: sketch_size, genome_size and the top-k screen ratio of n new genomes
void launch_stats_finalize(const uint32_t* active, const unsigned long long* ssum, const uint64_t* len,
                           uint32_t n, uint32_t* sketch_size, uint64_t* genome_size, float* ratio,
                           cudaStream_t st);
// pass B: the smallest (genome, bucket, probe) key of every still-zero Bloom byte writes it.
// claims[0 .. *n_claims): (seq_in_batch << h | bucket) of the buckets that registered a claim in pass A
void launch_bloom_commit(const unsigned long long* anc, const uint32_t* claims, const uint32_t* n_claims,
                         uint32_t n_seq, SketchParams p, uint8_t* bloom, uint32_t* owner, cudaStream_t st);
// fp[s][b] -> bit-plane rows (layout: scan.cu), genome column col0 + s
void launch_scatter_planes(const uint8_t* fp, uint32_t n_seq, int h, uint8_t* rows, uint64_t stride,
                           uint32_t col0, cudaStream_t st);
// bit-plane rows [row0, row0+nrows) <-> dense bucket-major bytes (the dump's layout)
void launch_planes_to_bytes(const uint8_t* rows, uint64_t stride, uint64_t row0, uint64_t nrows, uint32_t n,
                            uint8_t* out, cudaStream_t st);
void launch_bytes_to_planes(const uint8_t* in, uint64_t in_stride, uint64_t row0, uint64_t nrows, uint32_t n,
                            uint8_t* rows, uint64_t stride, cudaStream_t st);
// index merge: plane rows [row0, row0 + nrows) of another index (n_src genomes, staged at `src` with
// pitch src_stride) become columns col0 .. col0 + n_src - 1 of `rows`
void launch_merge_planes(uint8_t* rows, uint64_t stride, uint32_t col0, const uint8_t* src, uint64_t src_stride,
                         uint32_t n_src, uint64_t row0, uint64_t nrows, cudaStream_t st);
// dense read sketch -> (bucket << 8 | fp) list of buckets that are non-empty and pass Bloom
void launch_compact_list(const unsigned long long* anc, const uint8_t* fp, uint32_t n_seq,
                         SketchParams p, const uint8_t* bloom, const uint32_t* read_ids,
                         const uint64_t* list_off, uint32_t* list, uint32_t* list_len,
                         cudaStream_t st);
// Sparse path (reads with at most max_kmers k-mers): one CTA per read, smem hash table.
//   read_ids  : indices (into coff/len/list_off/list_len) of the reads to process
size_t sketch_reads_smem(uint64_t max_len, int k, uint32_t* slots_out);
int launch_sketch_reads(const uint8_t* chars, const uint64_t* coff, const uint64_t* len,
                        const uint32_t* read_ids, uint32_t n_ids, uint64_t max_len, SketchParams p,
                        const uint8_t* bloom, const uint64_t* list_off, uint32_t* list,
                        uint32_t* list_len, cudaStream_t st);
// stat[0] += sum list_len, stat[1] += sum list_len * n_genomes (device-side work counters)
void launch_account_rows(const uint32_t* list_len, uint32_t n, uint32_t n_genomes, unsigned long long* stat,
                         cudaStream_t st);
void launch_fill_u64(unsigned long long* p, uint64_t n, unsigned long long v, cudaStream_t st);
void launch_fill_u32(uint32_t* p, uint64_t n, uint32_t v, cudaStream_t st);
void launch_synth(uint8_t* chars, const uint64_t* coff, uint64_t seed, uint32_t first_g,
                  uint32_t n, uint64_t len, cudaStream_t st);
void launch_bloom_merge(uint8_t* dst, const uint8_t* src, uint64_t n, cudaStream_t st);

// ---- scan.cu --------------------------------------------------------------------
struct ScanPlan {
    int threads;        // consumer threads (multiple of 32)
    int J;              // 32-genome groups per thread
    uint32_t tile_w;    // genomes per tile (multiple of 32, <= 32 * threads * J)
    uint32_t n_tiles;   // genome tiles per row
    int stages;         // smem ring depth
    int blocks;         // 4-row carry-save blocks per stage (1 for wide tiles, up to 8)
    int whole_rows;     // 1: a stage row is a whole index row (one bulk copy), else two half copies
    uint32_t row_bytes; // pitch of a row inside a stage
    size_t smem;
    int grid;
    int sm_count;
    int narrow_gp;      // != 0: warp-per-read kernel for narrow shards, lanes per row
};
int scan_plan(uint32_t n_genomes, uint64_t stride, int sm_count, size_t smem_optin, int spare_sms, ScanPlan* out);
// counts[q * n_genomes + g] = #{entries e of read q : rows[bucket(e)][g] == fp(e)}
int launch_scan(const ScanPlan& plan, const uint8_t* rows, uint64_t stride, uint32_t n_genomes,
                const uint32_t* list, const uint64_t* list_off, const uint32_t* list_len,
                uint32_t n_reads, uint32_t* counts, uint32_t* work_counter, cudaStream_t st);

// ---- scan_tiled.cu ----------------------------------------------------------------
// The scan for read batches that cover the bucket space densely (long reads at small -h): a CTA
// owns a tile of reads x 1,024 genomes and streams every row of its genome tile through shared
// memory once (TMA tensor copies), each staged row serving all reads of the tile that hit it.
constexpr int TILED_MAX_H = 17;      // the list sort keeps one byte per bucket in shared memory
constexpr uint64_t TILED_SHORT_ENTRIES = 16380;   // lists up to here fit the 14 counter planes without a flush
struct TiledPlan {
    int J, warps;           // reads per consumer warp, consumer warps
    uint32_t tile_reads;    // J * warps
    uint32_t n_gt;          // genome tiles (<= 32 groups of 32 genomes each, balanced)
    uint32_t S;             // rows per ring stage (power of two)
    int stages;
    size_t smem;
    int grid;
};
int tiled_plan(uint32_t n_genomes, int h, int sm_count, size_t smem_optin_bytes, TiledPlan* out);
// the driver exports cuTensorMapEncodeTiled (looked up once, no link-time dependency on libcuda)
bool tiled_scan_available();
// entries a sorted list of `entries` entries occupies with its sentinels
uint32_t sorted_list_capacity(uint64_t entries);
// list (bucket << 8 | fp, any order) -> slist + soff[q] (8-byte entries {bucket * 1024, fp * 32}):
// ascending buckets, then sentinels.  soff[q] >= 128 and a multiple of 64: the first 128 entries of
// slist are a block of sentinels.
constexpr size_t SORTED_ENTRY_BYTES = 8;
int launch_sort_lists(const uint32_t* list, const uint64_t* list_off, const uint32_t* list_len, uint32_t n_reads, int h,
                       void* slist, const uint64_t* soff, cudaStream_t st);
int launch_scan_tiled(const TiledPlan& plan, const uint8_t* rows, uint64_t stride, uint32_t n_genomes, int h,
                      const void* slist, const uint64_t* soff, uint32_t n_reads, bool long_lists, uint32_t* counts,
                      uint32_t* work_counter, cudaStream_t st);

// ---- topk.cu --------------------------------------------------------------------
struct HitDev { uint32_t genome, matches; double jaccard, intersection; };
// heap: n_reads x nresults HitDev, len: n_reads.  chained across shards (first_id offsets).
void launch_topk(const uint32_t* counts, uint32_t n_reads, uint32_t n_genomes, uint32_t first_id,
                 const uint32_t* sketch_size, const uint64_t* genome_size, const float* ratio,
                 uint32_t nresults, uint32_t min_score, double min_intersection, HitDev* heap,
                 uint32_t* len, int finalize, cudaStream_t st);
// ratio[g] = float(genome_size[g] / sketch_size[g]): the top-k kernel's cheap screen

// ---- exact.cu -------------------------------------------------------------------
// open-addressing exact sets of canonical k-mers (utils.cpp:276 str2num); tables pre-set to ~0
void launch_exact_insert(const uint8_t* chars, const uint64_t* coff, const uint64_t* len,
                         uint32_t n_seq, uint64_t max_len, int k, unsigned long long* table,
                         uint64_t slots, unsigned long long* distinct, cudaStream_t st);
// per read r: set A in rtable[toff[r] .. toff[r+1]); inter[r] += |A n B|, distinctA[r] += |A|
void launch_exact_reads(const uint8_t* chars, const uint64_t* coff, const uint64_t* len,
                        uint32_t n_reads, uint64_t max_len, int k, unsigned long long* rtable,
                        const uint64_t* toff, const unsigned long long* tableB, uint64_t slotsB,
                        unsigned long long* inter, unsigned long long* distinctA, cudaStream_t st);

// the sort-merge variant of set B (MIEKKI_EXACT_SORT=1): all windows' k-mers at keys + koff[s],
// radix-sorted into `sorted`; *distinct += |B|
size_t exact_sort_temp_bytes(uint64_t n_keys, int k);
void launch_exact_sorted_build(const uint8_t* chars, const uint64_t* coff, const uint64_t* len, const uint64_t* koff,
                               uint32_t n_seq, uint64_t max_len, uint64_t n_keys, int k, unsigned long long* keys,
                               unsigned long long* sorted, void* temp, size_t temp_bytes, unsigned long long* distinct,
                               cudaStream_t st);
void launch_exact_reads_sorted(const uint8_t* chars, const uint64_t* coff, const uint64_t* len, uint32_t n_reads,
                               uint64_t max_len, int k, unsigned long long* rtable, const uint64_t* toff,
                               const unsigned long long* sortedB, uint64_t nB, unsigned long long* inter,
                               unsigned long long* distinctA, cudaStream_t st);

}  // namespace mk
