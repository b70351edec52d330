// sketch.cu -- partitioned min-hash sketch kernels (genomes and reads), sm_100a.
//
// Replaces Miekki::minhash_sketch_partition (Miekki.cpp:150-197), the per-genome part of
// insert_sequences (Miekki.cpp:287-311) and the Bloom filter (Miekki.cpp:121-146).
//
// Reference semantics reproduced here:
//   * n - k k-mers per sequence (loop "i + k < n", Miekki.cpp:162);
//   * per bucket the minimum fingerprint, and as `anc` the hash of the FIRST k-mer that
//     attains it (strict "<", Miekki.cpp:172): a 64-bit atomicMin on (fp << 56 | position);
//   * a fingerprint of 255 never registers (Miekki.cpp:172 with res[] preset to 255);
//   * Bloom bytes: the byte value belongs to the smallest (genome, bucket, probe) that maps
//     to it (order-independent restatement of Miekki.cpp:295-299).
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"
#include "kernels.h"

namespace mk {

// ---- character codes -------------------------------------------------------------
// bits 0-1 nuc2int (utils.cpp:31-49), bits 2-3 nuc2intrc (utils.cpp:107-125),
// bits 4-5 str2numstrand digit (utils.cpp:252-272), bit 6 "valid for str2numstrand".
__device__ __forceinline__ uint32_t char_code(uint32_t c) {
    uint32_t f = c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 0u;
    uint32_t r = c == 'A' ? 3u : c == 'C' ? 2u : c == 'G' ? 1u : 0u;
    uint32_t u = c & 0xDFu;  // fold case for the four letters only
    bool letter = (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z');
    uint32_t pc = 0, pv = 0;
    if (letter && (u == 'A' || u == 'C' || u == 'G' || u == 'T')) {
        pv = 1;
        pc = u == 'C' ? 1u : u == 'G' ? 2u : u == 'T' ? 3u : 0u;
    }
    return f | (r << 2) | (pc << 4) | (pv << 6);
}

__device__ __forceinline__ void fill_lut(uint8_t* lut) {
    for (int c = threadIdx.x; c < 256; c += blockDim.x) lut[c] = (uint8_t)char_code((uint32_t)c);
}

// all of the first min(k-1, n) characters valid for str2numstrand?  (else the prefix is 0)
__device__ __forceinline__ bool prefix_valid(const uint8_t* __restrict__ s, uint64_t n, int k,
                                             const uint8_t* lut) {
    const int m = (uint64_t)(k - 1) < n ? (k - 1) : (int)n;
    bool ok = true;
    for (int i = 0; i < m; ++i) ok = ok && ((lut[s[i]] >> 6) & 1u);
    return ok;
}

// encode characters [16w, 16w+16) of a sequence (given as 16 raw bytes) into one F and one R word
__device__ __forceinline__ void encode_word_raw(const uint4 raw, uint64_t n, uint64_t w, int k, bool pvalid,
                                                const uint8_t* lut, uint32_t& F, uint32_t& R) {
    const uint32_t q[4] = {raw.x, raw.y, raw.z, raw.w};
    uint32_t f = 0, r = 0;
    #pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint64_t pos = 16 * w + i;
        const uint32_t c = (q[i >> 2] >> (8 * (i & 3))) & 0xFFu;
        const uint32_t code = lut[c];
        uint32_t fd, rd;
        if (pos < (uint64_t)(k - 1)) {           // prefix digits: str2numstrand, then rcb
            fd = pvalid ? ((code >> 4) & 3u) : 0u;
            rd = 3u - fd;
        } else {
            fd = code & 3u;
            rd = (code >> 2) & 3u;
        }
        if (pos >= n) { fd = 0; rd = 0; }
        f |= fd << (30 - 2 * i);
        r |= rd << (2 * i);
    }
    F = f;
    R = r;
}
__device__ __forceinline__ void encode_word(const uint8_t* __restrict__ s, uint64_t n, uint64_t w,
                                            int k, bool pvalid, const uint8_t* lut, uint32_t& F,
                                            uint32_t& R) {
    encode_word_raw(*reinterpret_cast<const uint4*>(s + 16 * w), n, w, k, pvalid, lut, F, R);  // 16-byte aligned start
}

// Fast path of encode_word: all 16 characters are upper-case A/C/G/T and none lies in the
// prefix zone, so both encoders agree (F digit = A0 C1 G2 T3, R digit = 3 - F digit) and the
// word is packed arithmetically, four characters per 32-bit lane.  Returns false (F, R
// untouched) when any other byte is present: the caller then takes the table path.
__device__ __forceinline__ bool encode_word_acgt(const uint4 raw, uint32_t& F, uint32_t& R) {
    const uint32_t q[4] = {raw.x, raw.y, raw.z, raw.w};
    uint32_t bad = 0, f = 0;
    #pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t x = (q[i] >> 1) & 0x03030303u;               // A0 C1 G3 T2
        const uint32_t t = (x >> 1) & ~x & 0x01010101u;             // 1 per 'T'
        const uint32_t expect = (0x41414141u | (x << 1)) ^ (t * 0x11u);   // 41 43 47 | 45^11 = 54
        bad |= q[i] ^ expect;
        const uint32_t code = x ^ ((x >> 1) & 0x01010101u);         // A0 C1 G2 T3
        // bytes c0..c3 (c0 = first character) -> c0<<6 | c1<<4 | c2<<2 | c3 in the top byte
        f |= ((code * 0x40100401u) >> 24) << (24 - 8 * i);
    }
    if (bad) return false;
    F = f;
    // R digit of base j (bits 2j) = 3 - F digit of base j (bits 30-2j): reverse the digit order
    uint32_t r = __brev(f);
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    R = ~r;
    return true;
}

// ---- dense path ------------------------------------------------------------------

__global__ void __launch_bounds__(256)
encode_planes_kernel(const uint8_t* __restrict__ chars, const uint64_t* __restrict__ coff,
                     const uint64_t* __restrict__ len, const uint64_t* __restrict__ woff, int k,
                     uint32_t* __restrict__ planeF, uint32_t* __restrict__ planeR) {
    __shared__ uint8_t lut[256];
    fill_lut(lut);
    __syncthreads();
    const uint32_t s = blockIdx.y;
    const uint64_t n = len[s];
    const uint64_t nw = (n + 15) / 16 + 2;       // two zero words of padding
    const uint8_t* seq = chars + coff[s];
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nw;
         w += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t F = 0, R = 0;
        if (16 * w < n) {
            const bool plain = 16 * w >= (uint64_t)(k - 1) && 16 * w + 16 <= n;
            if (!plain || !encode_word_acgt(*reinterpret_cast<const uint4*>(seq + 16 * w), F, R)) {
                const bool pv = (16 * w < (uint64_t)(k - 1)) ? prefix_valid(seq, n, k, lut) : true;
                encode_word(seq, n, w, k, pv, lut, F, R);
            }
        }
        // the two planes are interleaved word by word (planeR == planeF + 1): the three words a
        // k-mer needs from each plane are then 24 contiguous bytes, i.e. one or two sectors
        // instead of two to four for the random look-ups of resolve_kernel
        *reinterpret_cast<uint2*>(planeF + 2 * (woff[s] + w)) = make_uint2(F, R);
        (void)planeR;
    }
}

// The same planes from sequences packed on the host (pack.h): a packed word holds the forward
// digits of 16 upper-case A/C/G/T outside the prefix, so F is the word itself and R its digit-wise
// complement in reverse order; every other word (prefix, ragged end, N, lower case ...) arrives
// as 16 raw bytes in the exception list and goes through the general encoder afterwards.
__global__ void __launch_bounds__(256)
expand_planes_kernel(const uint32_t* __restrict__ packed, const uint64_t* __restrict__ len,
                     const uint64_t* __restrict__ woff, uint32_t* __restrict__ planes) {
    const uint32_t s = blockIdx.y;
    const uint64_t n = len[s];
    const uint64_t nw = (n + 15) / 16 + 2;       // two zero words of padding
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nw;
         w += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t F = 0, R = 0;
        if (16 * w + 16 <= n) {                  // whole words only; the others are exceptions or padding
            F = packed[woff[s] + w];
            uint32_t r = __brev(F);
            r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
            R = ~r;
        }
        *reinterpret_cast<uint2*>(planes + 2 * (woff[s] + w)) = make_uint2(F, R);
    }
}
__global__ void __launch_bounds__(256)
patch_planes_kernel(const PackExcDev* __restrict__ exc, uint32_t n_exc, uint32_t seq0, const uint64_t* __restrict__ len,
                    const uint64_t* __restrict__ woff, const uint8_t* __restrict__ pvalid, int k,
                    uint32_t* __restrict__ planes) {
    __shared__ uint8_t lut[256];
    fill_lut(lut);
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_exc) return;
    const PackExcDev e = exc[i];
    const uint32_t s = e.seq - seq0;
    uint32_t F, R;
    encode_word_raw(e.bytes, len[s], e.word, k, pvalid[s] != 0, lut, F, R);
    *reinterpret_cast<uint2*>(planes + 2 * (woff[s] + e.word)) = make_uint2(F, R);
}

// one thread = 16 consecutive k-mer start positions (one plane word + two neighbours)
// key = fp << 56 | position << ks | tag.  ks = KEY_TAG_BITS when every position fits 32 bits:
// the tag is then the top KEY_TAG_BITS bits of the canonical k-mer, which lets resolve_kernel
// decide "every Bloom byte this k-mer can touch is already set" without looking the k-mer up
// again.  The order of keys is still (fp, position): the tag sits below both.
__global__ void __launch_bounds__(256)
sketch_dense_kernel(const uint32_t* __restrict__ planeF, const uint32_t* __restrict__ planeR,
                    const uint64_t* __restrict__ len, const uint64_t* __restrict__ woff, int k, int h,
                    unsigned long long* __restrict__ keys, int prefilter, int ks) {
    const uint32_t s = blockIdx.y;
    const uint64_t n = len[s];
    if (n <= (uint64_t)k) return;
    const uint64_t nk = n - k;                    // quirk G1: the last k-mer is not visited
    const uint64_t nwk = (nk + 15) / 16;
    const uint64_t kmask = (1ull << (2 * k)) - 1;
    const uint64_t pmask = (1ull << (64 - h)) - 1;
    const int tag_shift = key_tag_shift(k);
    const uint2* P = reinterpret_cast<const uint2*>(planeF) + woff[s];   // .x = F word, .y = R word
    (void)planeR;
    unsigned long long* kz = keys + ((uint64_t)s << h);
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwk;
         w += (uint64_t)gridDim.x * blockDim.x) {
        const uint2 p0 = P[w], p1 = P[w + 1], p2 = P[w + 2];
        const uint32_t f0 = p0.x, f1 = p1.x, f2 = p2.x;
        const uint32_t r0 = p0.y, r1 = p1.y, r2 = p2.y;
        #pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint64_t pos = 16 * w + j;
            if (pos < nk) {
                const uint64_t canon = kmer_canon(f0, f1, f2, r0, r1, r2, j, k, kmask);
                const uint64_t x = revhash64(canon);
                const uint32_t fp = mantis(x & pmask, h);
                if (fp != EMPTY_FP) {
                    unsigned long long* slot = kz + (x >> (64 - h));
                    unsigned long long key = ((unsigned long long)fp << POS_BITS) | (pos << ks);
                    if (ks) key |= canon >> tag_shift;
                    // Keys only ever decrease, so a (possibly stale) read that is already <= key
                    // proves the atomic would change nothing: most k-mers of a bucket lose.
                    if (!prefilter || key < *reinterpret_cast<volatile unsigned long long*>(slot)) atomicMin(slot, key);
                }
            }
        }
    }
}

// key = (seq_in_batch, bucket, probe) packed so that u32 order == lexicographic order
__device__ __forceinline__ uint32_t owner_key(uint32_t s, uint32_t bucket, uint32_t i, int h) {
    return (((s << h) | bucket) << 3) | i;
}

// The kernel is a chain of three dependent random look-ups per bucket (key -> plane words at
// the winning position -> Bloom byte) and was purely latency bound with one bucket per thread
// (ncu: 75 % long-scoreboard stalls, DRAM 8 %).  Each thread therefore walks RES_U buckets in
// lock step, stage by stage, so that RES_U independent loads are in flight per stage.
//
// Build fast path (tagged keys, ks != 0): the reference's Bloom table only ever gains bytes
// (Miekki.cpp:127-128 writes into zero bytes only), and for the default k = 31, b = 33 it has
// just 2^26 reachable bytes, so after a few hundred genomes nearly every insert is a no-op.
// `pair_full` has one bit per 4 KB page of the table: "this page and the next hold no zero
// byte" (bloom_pages_kernel, recomputed before every launch).  The key's tag bounds the
// canonical k-mer to a range whose five probes stay inside those two pages, so a set bit
// proves the insert changes nothing: `anc` is not needed and all three look-ups are skipped.
constexpr int RES_U = 4;
constexpr uint32_t NO_CLAIM = 0xFFFFFFFFu;     // claim-list entry struck off by resolve_slow_kernel
constexpr uint32_t PAIR_WORDS_MAX = 1032;      // (2^27 + 16 window bytes) / 4096 / 32, rounded up

__global__ void __launch_bounds__(256, 4)
resolve_kernel(unsigned long long* __restrict__ keys_anc, const uint32_t* __restrict__ planeF,
               const uint64_t* __restrict__ woff, SketchParams p, uint8_t* __restrict__ fp_out,
               uint32_t* __restrict__ active, unsigned long long* __restrict__ ssum,
               const uint8_t* __restrict__ bloom, uint32_t* __restrict__ owner,
               const uint32_t* __restrict__ pair_full, uint32_t pair_words, int ks,
               uint32_t* __restrict__ claims, uint32_t* __restrict__ n_claims) {
    __shared__ uint32_t s_pair[PAIR_WORDS_MAX];
    __shared__ uint32_t s_act[8], s_claims[8], s_base;
    __shared__ unsigned long long s_sum[8];
    const bool tagged = owner != nullptr && pair_full != nullptr && ks != 0;
    if (tagged) {
        for (uint32_t i = threadIdx.x; i < pair_words; i += blockDim.x) s_pair[i] = pair_full[i];
        __syncthreads();
    }
    const uint32_t s = blockIdx.y;
    const uint32_t B = 1u << p.h;
    const uint2* P = reinterpret_cast<const uint2*>(planeF) + woff[s];
    const uint64_t row = (uint64_t)s << p.h;
    const uint64_t kmask = (1ull << (2 * p.k)) - 1;
    const int tag_shift = key_tag_shift(p.k);
    const int page_shift = (int)p.bloom_log2 + 3 + BLOOM_PAGE_LOG2;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t act = 0;
    unsigned long long sum = 0;

    // A CTA walks several tiles of 256 * RES_U buckets: the page bitmap is loaded once, and a
    // short-lived CTA per tile spent most of its life in the launch / drain latency chain.
    const uint32_t n_tiles = (B + 256 * RES_U - 1) / (256 * RES_U);
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint32_t b[RES_U];
        unsigned long long key[RES_U];
        bool slow[RES_U];                                       // needs anc (and maybe a Bloom claim)
        #pragma unroll
        for (int u = 0; u < RES_U; ++u) {                       // stage 1: keys (coalesced)
            b[u] = (tile * RES_U + u) * blockDim.x + threadIdx.x;
            key[u] = b[u] < B ? keys_anc[row + b[u]] : EMPTY_KEY;
            slow[u] = (key[u] >> POS_BITS) != EMPTY_FP;
            if (tagged && slow[u]) {
                const uint32_t page = (uint32_t)(((key[u] & KEY_TAG_MASK) << tag_shift) >> page_shift);
                if ((s_pair[page >> 5] >> (page & 31)) & 1u) slow[u] = false;
            }
        }
        uint2 w0[RES_U], w1[RES_U], w2[RES_U];
        #pragma unroll
        for (int u = 0; u < RES_U; ++u) {                       // stage 2: plane words at the winning position
            w0[u] = w1[u] = w2[u] = make_uint2(0u, 0u);
            if (slow[u]) {
                const uint64_t w = ((key[u] & POS_MASK) >> ks) >> 4;
                w0[u] = P[w];
                w1[u] = P[w + 1];
                w2[u] = P[w + 2];
            }
        }
        unsigned long long anc[RES_U];
        uint64_t slot[RES_U][2];
        uint32_t probe[RES_U][2];
        int nd[RES_U];
        uint8_t cell[RES_U][2];
        #pragma unroll
        for (int u = 0; u < RES_U; ++u) {                       // stage 3: hashes, Bloom bytes
            anc[u] = EMPTY_ANC;
            nd[u] = 0;
            cell[u][0] = cell[u][1] = 1;
            if (slow[u]) {
                const int j = (int)(((key[u] & POS_MASK) >> ks) & 15);
                anc[u] = kmer_hash(w0[u].x, w1[u].x, w2[u].x, w0[u].y, w1[u].y, w2[u].y, j, p.k, kmask);
                if (owner != nullptr) {
                    BloomProbe pr(anc[u]);
                    nd[u] = bloom_first_probes(pr, p.bloom_log2, slot[u], probe[u]);
                    #pragma unroll
                    for (int q = 0; q < 2; ++q)
                        if (q < nd[u]) {
                            const uint64_t byte = slot[u][q] >> 3;
                            cell[u][q] = byte < p.bloom_window ? bloom[byte] : (uint8_t)1;
                        }
                }
            }
        }
        uint32_t claim_bits = 0;
        #pragma unroll
        for (int u = 0; u < RES_U; ++u) {                       // stage 4: claims and stores
            const bool in = b[u] < B;
            const uint32_t fp = (uint32_t)(key[u] >> POS_BITS);
            bool claimed = false;
            if (in && fp != EMPTY_FP) {
                act += 1;
                sum += 1ull << (31 - (fp >> 3));                // 2^-(fp>>3) in units of 2^-31 (Miekki.cpp:293)
                if (owner != nullptr) {                         // Bloom pass A (Miekki.cpp:295-299)
                    #pragma unroll
                    for (int q = 0; q < 2; ++q)
                        if (q < nd[u] && cell[u][q] == 0) {
                            atomicMin(owner + (slot[u][q] >> 3), owner_key(s, b[u], probe[u][q], p.h));
                            claimed = true;
                        }
                }
            }
            claim_bits |= claimed ? (1u << u) : 0u;
            if (!in) continue;
            // during a build only claimants are looked at again (bloom_commit_kernel reads their anc)
            if (owner == nullptr || slow[u]) keys_anc[row + b[u]] = anc[u];
            fp_out[row + b[u]] = (uint8_t)fp;
        }
        if (owner == nullptr) continue;
        // the claimant list (pass B only revisits claimants): one global atomic per tile, and
        // none at all once the Bloom table is saturated
        if (!__syncthreads_or((int)claim_bits)) continue;
        const uint32_t mine = __popc(claim_bits);
        uint32_t before = mine;                                 // inclusive prefix sum over the warp
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, before, o);
            if (lane >= (uint32_t)o) before += t;
        }
        if (lane == 31) s_claims[warp] = before;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t cl = 0;
            for (unsigned w = 0; w < blockDim.x / 32; ++w) {
                const uint32_t x = s_claims[w];
                s_claims[w] = cl;                               // -> exclusive offset of the warp
                cl += x;
            }
            s_base = atomicAdd(n_claims, cl);
        }
        __syncthreads();
        if (mine) {
            uint32_t at = s_base + s_claims[warp] + before - mine;
            #pragma unroll
            for (int u = 0; u < RES_U; ++u)
                if (claim_bits & (1u << u)) claims[at++] = (s << p.h) | b[u];
        }
        __syncthreads();                                        // s_claims / s_base are reused by the next tile
    }
    // per-sequence statistics: warp shuffles, then one pair of global atomics per block
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        act += __shfl_xor_sync(0xffffffffu, act, o);
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    if (lane == 0) {
        s_act[warp] = act;
        s_sum[warp] = sum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t a = 0;
        unsigned long long t = 0;
        for (unsigned w = 0; w < blockDim.x / 32; ++w) { a += s_act[w]; t += s_sum[w]; }
        if (a) {
            atomicAdd(active + s, a);
            atomicAdd(ssum + s, t);
        }
    }
}

// ---- build with tagged keys: the same work as resolve_kernel, split in two ------------------
// resolve_fast_kernel streams every key once (32 bytes per thread and tile, two tiles in
// flight): fingerprint bytes, per-genome statistics, and the list of buckets whose k-mer may
// still change the Bloom table (page bitmap says "not saturated").  resolve_slow_kernel then
// looks only those buckets up (key -> plane words -> anc -> Bloom bytes), registers claims and
// keeps the claimants in the list for bloom_commit_kernel.  Once the table is saturated the
// list is almost empty and the build no longer does any random look-up per bucket.
__global__ void __launch_bounds__(256)
resolve_fast_kernel(const unsigned long long* __restrict__ keys, SketchParams p, uint8_t* __restrict__ fp_out,
                    uint32_t* __restrict__ active, unsigned long long* __restrict__ ssum,
                    const uint32_t* __restrict__ pair_full, uint32_t pair_words,
                    uint32_t* __restrict__ slow_list, uint32_t* __restrict__ n_slow) {
    __shared__ uint32_t s_pair[PAIR_WORDS_MAX];
    __shared__ uint32_t s_act[8], s_cnt[8], s_base;
    __shared__ unsigned long long s_sum[8];
    for (uint32_t i = threadIdx.x; i < pair_words; i += blockDim.x) s_pair[i] = pair_full[i];
    __syncthreads();
    const uint32_t s = blockIdx.y;
    const uint32_t B = 1u << p.h;                               // multiple of 1024 (host: h >= 10)
    const uint64_t row = (uint64_t)s << p.h;
    // page_shift (>= 47) always exceeds tag_shift (<= 38)
    const int tag_to_page = (int)p.bloom_log2 + 3 + BLOOM_PAGE_LOG2 - key_tag_shift(p.k);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_tiles = B >> 10;
    uint32_t act = 0;
    unsigned long long sum = 0;
    constexpr int T = 2;                                        // tiles in flight per CTA
    for (uint32_t tile0 = blockIdx.x * T; tile0 < n_tiles; tile0 += gridDim.x * T) {
        ulonglong2 k01[T], k23[T];
        #pragma unroll
        for (int t = 0; t < T; ++t) {
            const uint32_t b4 = (tile0 + t < n_tiles) ? ((tile0 + t) << 10) + 4 * threadIdx.x : 0xFFFFFFFFu;
            if (b4 != 0xFFFFFFFFu) {
                const ulonglong2* src = reinterpret_cast<const ulonglong2*>(keys + row + b4);
                k01[t] = src[0];
                k23[t] = src[1];
            } else {
                k01[t] = k23[t] = make_ulonglong2(EMPTY_KEY, EMPTY_KEY);
            }
        }
        #pragma unroll
        for (int t = 0; t < T; ++t) {
            if (tile0 + t >= n_tiles) break;                    // uniform over the CTA
            const uint32_t b4 = ((tile0 + t) << 10) + 4 * threadIdx.x;
            const unsigned long long key[4] = {k01[t].x, k01[t].y, k23[t].x, k23[t].y};
            uint32_t packed = 0, slow_bits = 0;
            uint32_t part = 0;                                  // <= 4 * 2^31 overflows 32 bits: folded below
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t fp = (uint32_t)(key[u] >> POS_BITS);
                packed |= fp << (8 * u);
                if (fp != EMPTY_FP) {
                    act += 1;
                    const uint32_t term = 1u << (31 - (fp >> 3));  // 2^-(fp>>3) in units of 2^-31 (Miekki.cpp:293)
                    sum += part + term < part ? (1ull << 32) : 0ull;
                    part += term;
                    // page of the k-mer range the tag stands for: (tag << tag_shift) >> page_shift
                    const uint32_t tag = (uint32_t)key[u] & (uint32_t)KEY_TAG_MASK;
                    const uint32_t page = tag_to_page < 32 ? tag >> tag_to_page : 0u;
                    if (!((s_pair[page >> 5] >> (page & 31)) & 1u)) slow_bits |= 1u << u;
                }
            }
            sum += part;
            *reinterpret_cast<uint32_t*>(fp_out + row + b4) = packed;
            if (!__syncthreads_or((int)slow_bits)) continue;
            const uint32_t mine = __popc(slow_bits);
            uint32_t before = mine;                             // inclusive prefix sum over the warp
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t x = __shfl_up_sync(0xffffffffu, before, o);
                if (lane >= (uint32_t)o) before += x;
            }
            if (lane == 31) s_cnt[warp] = before;
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t cl = 0;
                for (unsigned w = 0; w < blockDim.x / 32; ++w) {
                    const uint32_t x = s_cnt[w];
                    s_cnt[w] = cl;
                    cl += x;
                }
                s_base = atomicAdd(n_slow, cl);
            }
            __syncthreads();
            if (mine) {
                uint32_t at = s_base + s_cnt[warp] + before - mine;
                #pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (slow_bits & (1u << u)) slow_list[at++] = (s << p.h) | (b4 + u);
            }
            __syncthreads();
        }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        act += __shfl_xor_sync(0xffffffffu, act, o);
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    if (lane == 0) {
        s_act[warp] = act;
        s_sum[warp] = sum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t a = 0;
        unsigned long long t = 0;
        for (unsigned w = 0; w < blockDim.x / 32; ++w) { a += s_act[w]; t += s_sum[w]; }
        if (a) {
            atomicAdd(active + s, a);
            atomicAdd(ssum + s, t);
        }
    }
}

__global__ void __launch_bounds__(256)
resolve_slow_kernel(unsigned long long* __restrict__ keys_anc, const uint32_t* __restrict__ planeF,
                    const uint64_t* __restrict__ woff, SketchParams p, const uint8_t* __restrict__ bloom,
                    uint32_t* __restrict__ owner, uint32_t* __restrict__ list,
                    const uint32_t* __restrict__ n_list, int ks) {
    const uint32_t total = *n_list;
    const uint32_t bmask = (1u << p.h) - 1;
    const uint64_t kmask = (1ull << (2 * p.k)) - 1;
    const uint32_t step = gridDim.x * blockDim.x;
    for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += RES_U * step) {
        uint32_t id[RES_U];
        bool live[RES_U];
        unsigned long long key[RES_U];
        #pragma unroll
        for (int u = 0; u < RES_U; ++u) {                       // stage 0: list entries (coalesced)
            const uint32_t i = i0 + u * step;
            live[u] = i < total;
            id[u] = live[u] ? list[i] : 0u;
        }
        #pragma unroll
        for (int u = 0; u < RES_U; ++u) key[u] = live[u] ? keys_anc[id[u]] : 0ull;     // stage 1
        uint2 w0[RES_U], w1[RES_U], w2[RES_U];
        #pragma unroll
        for (int u = 0; u < RES_U; ++u) {                       // stage 2: plane words at the winning position
            w0[u] = w1[u] = w2[u] = make_uint2(0u, 0u);
            if (live[u]) {
                const uint2* P = reinterpret_cast<const uint2*>(planeF) + woff[id[u] >> p.h];
                const uint64_t w = ((key[u] & POS_MASK) >> ks) >> 4;
                w0[u] = P[w];
                w1[u] = P[w + 1];
                w2[u] = P[w + 2];
            }
        }
        uint64_t slot[RES_U][2];
        uint32_t probe[RES_U][2];
        int nd[RES_U];
        uint8_t cell[RES_U][2];
        #pragma unroll
        for (int u = 0; u < RES_U; ++u) {                       // stage 3: anc, Bloom bytes
            nd[u] = 0;
            cell[u][0] = cell[u][1] = 1;
            if (live[u]) {
                const int j = (int)(((key[u] & POS_MASK) >> ks) & 15);
                const unsigned long long anc =
                    kmer_hash(w0[u].x, w1[u].x, w2[u].x, w0[u].y, w1[u].y, w2[u].y, j, p.k, kmask);
                keys_anc[id[u]] = anc;                          // bloom_commit_kernel reads it
                BloomProbe pr(anc);
                nd[u] = bloom_first_probes(pr, p.bloom_log2, slot[u], probe[u]);
                #pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (q < nd[u]) {
                        const uint64_t byte = slot[u][q] >> 3;
                        cell[u][q] = byte < p.bloom_window ? bloom[byte] : (uint8_t)1;
                    }
            }
        }
        #pragma unroll
        for (int u = 0; u < RES_U; ++u) {                       // stage 4: claims (Miekki.cpp:295-299, pass A)
            if (!live[u]) continue;
            bool claimed = false;
            #pragma unroll
            for (int q = 0; q < 2; ++q)
                if (q < nd[u] && cell[u][q] == 0) {
                    atomicMin(owner + (slot[u][q] >> 3), owner_key(id[u] >> p.h, id[u] & bmask, probe[u][q], p.h));
                    claimed = true;
                }
            if (!claimed) list[i0 + u * step] = NO_CLAIM;       // pass B skips it
        }
    }
}

// Which 4 KB pages of the Bloom table hold no zero byte?  One warp per page -> full8[page];
// then pair bit j = full8[j] && full8[j + 1] (a k-mer's probes may cross into the next page).
// Only bytes below `reach` can ever be probed (api.cu: mk_bloom_reach), so bytes from there on do
// not count, and a page that starts at or past `reach` is vacuously full.
__global__ void __launch_bounds__(256)
bloom_pages_kernel(const uint8_t* __restrict__ bloom, uint64_t reach, uint32_t n_pages,
                   uint8_t* __restrict__ full8) {
    const uint32_t page = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (page >= n_pages) return;
    const uint64_t base = (uint64_t)page << BLOOM_PAGE_LOG2;
    uint32_t zero = 0;
    #pragma unroll
    for (int i = 0; i < (1 << BLOOM_PAGE_LOG2) / 16 / 32; ++i) {
        const uint64_t off = base + 16ull * (i * 32 + lane);
        if (off + 16 <= reach) {
            const uint4 v = *reinterpret_cast<const uint4*>(bloom + off);
            zero |= (v.x - 0x01010101u) & ~v.x;
            zero |= (v.y - 0x01010101u) & ~v.y;
            zero |= (v.z - 0x01010101u) & ~v.z;
            zero |= (v.w - 0x01010101u) & ~v.w;
        } else {
            for (uint64_t o = off; o < reach; ++o) zero |= bloom[o] == 0 ? 0x80u : 0u;
        }
    }
    const bool ok = __all_sync(0xffffffffu, (zero & 0x80808080u) == 0);
    if (lane == 0) full8[page] = ok ? 1 : 0;
}
__global__ void __launch_bounds__(256)
bloom_pairs_kernel(const uint8_t* __restrict__ full8, uint32_t n_pages, uint32_t n_words,
                   uint32_t* __restrict__ pair_full) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t bits = 0;
    for (uint32_t j = 0; j < 32; ++j) {
        const uint32_t pg = 32 * w + j;
        // no page after the last one: nothing there can be probed
        if (pg < n_pages && full8[pg] && (pg + 1 >= n_pages || full8[pg + 1])) bits |= 1u << j;
    }
    pair_full[w] = bits;
}

// Miekki.cpp:303-311 from the exact integer statistics: S = ssum31 * 2^-31 is exact, the
// u32 product wraps (quirk G2), and mul / div are single IEEE roundings like the host code.
__global__ void __launch_bounds__(256)
stats_finalize_kernel(const uint32_t* __restrict__ active, const unsigned long long* __restrict__ ssum,
                      const uint64_t* __restrict__ len, uint32_t n, uint32_t* __restrict__ sketch_size,
                      uint64_t* __restrict__ genome_size, float* __restrict__ ratio) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t a = active[i];
    const double S = __dmul_rn(__ull2double_rn(ssum[i]), 4.656612873077392578125e-10);   // 2^-31
    const uint32_t sq = a * a;
    const double card = __ddiv_rn(__dmul_rn(0.72134, (double)sq), S);
    const double dl = __ull2double_rn(len[i]);
    uint64_t gs;
    if (card > dl) gs = len[i];
    else if (card != card) gs = 0x8000000000000000ull;      // 0/0: what cvttsd2si yields on x86
    else gs = (uint64_t)card;
    sketch_size[i] = a;
    genome_size[i] = gs;
    ratio[i] = (float)__ddiv_rn(__ull2double_rn(gs), (double)a);
}

__global__ void __launch_bounds__(256)
bloom_commit_kernel(const unsigned long long* __restrict__ anc, const uint32_t* __restrict__ claims,
                    const uint32_t* __restrict__ n_claims, SketchParams p, uint8_t* __restrict__ bloom,
                    uint32_t* __restrict__ owner) {
    // RES_U claimants per thread, stage by stage, like resolve_kernel (random owner look-ups)
    const uint32_t total = *n_claims;
    const uint32_t bmask = (1u << p.h) - 1;
    const uint32_t step = gridDim.x * blockDim.x;
    for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += RES_U * step) {
        uint32_t id[RES_U];
        bool live[RES_U];
        unsigned long long a[RES_U];
        #pragma unroll
        for (int u = 0; u < RES_U; ++u) {
            const uint32_t i = i0 + u * step;
            id[u] = i < total ? claims[i] : NO_CLAIM;
            live[u] = id[u] != NO_CLAIM;                        // resolve_slow_kernel struck it off
        }
        #pragma unroll
        for (int u = 0; u < RES_U; ++u) a[u] = live[u] ? anc[id[u]] : 0ull;
        uint64_t slot[RES_U][2];
        uint32_t probe[RES_U][2], own[RES_U][2];
        int nd[RES_U];
        #pragma unroll
        for (int u = 0; u < RES_U; ++u) {
            nd[u] = 0;
            if (live[u]) {
                BloomProbe pr(a[u]);
                nd[u] = bloom_first_probes(pr, p.bloom_log2, slot[u], probe[u]);
                #pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (q < nd[u]) {
                        const uint64_t byte = slot[u][q] >> 3;
                        own[u][q] = byte < p.bloom_window ? owner[byte] : 0xFFFFFFFFu;
                    }
            }
        }
        #pragma unroll
        for (int u = 0; u < RES_U; ++u)
            #pragma unroll
            for (int q = 0; q < 2; ++q)
                if (q < nd[u] && own[u][q] == owner_key(id[u] >> p.h, id[u] & bmask, probe[u][q], p.h)) {
                    const uint64_t byte = slot[u][q] >> 3;
                    bloom[byte] = (uint8_t)(1u << (slot[u][q] & 7));   // Miekki.cpp:128
                    owner[byte] = 0xFFFFFFFFu;                          // the winner also clears its claim
                }
    }
}

// ---- bit-plane row layout (see scan.cu) ------------------------------------------------
// row b = [half 0 | half 1] (stride/2 bytes each); group j = genomes 32j..32j+31 keeps planes
// 0-3 as one uint4 at half0 + 16 j and planes 4-7 at half1 + 16 j; bit i of plane p = bit p of
// the fingerprint of genome 32 j + i.  An untouched row is all ones (every fingerprint 255).
__device__ __forceinline__ uint4* plane_ptr(uint8_t* rows, uint64_t stride, uint64_t b, uint32_t group,
                                            int half) {
    return reinterpret_cast<uint4*>(rows + b * stride + (half ? (stride >> 1) : 0) + 16ull * group);
}

// fp[s][b] (s = genome col0 + s) -> plane bits.  One thread per (bucket, 32-genome group the
// chunk touches): the <= 32 fingerprint loads are independent (contiguous along b across the
// warp) and fully unrolled.  A group whose first genome belongs to this chunk has never been
// written (genomes are appended in id order, untouched cells are all ones), so it is stored
// without reading the row; only a chunk that starts inside a group merges with what is there.
__global__ void __launch_bounds__(256)
scatter_planes_kernel(const uint8_t* __restrict__ fp, uint32_t n_seq, int h, uint8_t* __restrict__ rows,
                      uint64_t stride, uint32_t col0) {
    // neighbouring lanes own the two groups of one 32-byte sector of a row half
    const uint64_t B = 1ull << h;
    const uint64_t b = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    const uint32_t grp = (col0 >> 5) + 2 * blockIdx.y + (threadIdx.x & 1u);
    if (b >= B || grp > ((col0 + n_seq - 1) >> 5)) return;
    const uint32_t lo = max(col0, grp << 5), hi = min(col0 + n_seq, (grp + 1) << 5);
    const uint32_t bit0 = lo & 31u, cnt = hi - lo;
    const uint8_t* src = fp + ((uint64_t)(lo - col0) << h) + b;
    uint32_t f[32];
    #pragma unroll
    for (uint32_t i = 0; i < 32; ++i) f[i] = i < cnt ? src[(uint64_t)i << h] : 0xFFu;
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    #pragma unroll
    for (uint32_t i = 0; i < 32; ++i) {
        #pragma unroll
        for (int p = 0; p < 8; ++p) w[p] |= ((f[i] >> p) & 1u) << i;
    }
    // bits [bit0, bit0 + cnt) of the group are ours; slots past cnt were filled with 255
    const uint32_t keep = bit0 ? ((1u << bit0) - 1u) : 0u;      // genomes of earlier chunks
    #pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint4* q = plane_ptr(rows, stride, b, grp, half);
        uint4 v = make_uint4(w[4 * half + 0] << bit0, w[4 * half + 1] << bit0, w[4 * half + 2] << bit0,
                             w[4 * half + 3] << bit0);
        if (keep) {
            const uint4 old = *q;
            v.x |= old.x & keep;
            v.y |= old.y & keep;
            v.z |= old.z & keep;
            v.w |= old.w & keep;
        }
        *q = v;
    }
}

// plane layout -> dense bytes out[(b - row0) * n + g] for rows [row0, row0 + nrows) (dump)
__global__ void __launch_bounds__(256)
planes_to_bytes_kernel(const uint8_t* __restrict__ rows, uint64_t stride, uint64_t row0, uint64_t nrows,
                       uint32_t n, uint8_t* __restrict__ out) {
    const uint32_t groups = (n + 31) / 32;
    const uint64_t total = nrows * groups;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / groups;
        const uint32_t grp = (uint32_t)(i % groups);
        const uint4 a = *plane_ptr(const_cast<uint8_t*>(rows), stride, row0 + r, grp, 0);
        const uint4 c = *plane_ptr(const_cast<uint8_t*>(rows), stride, row0 + r, grp, 1);
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
        const uint32_t cnt = min(32u, n - grp * 32);
        uint8_t* o = out + r * n + (uint64_t)grp * 32;
        for (uint32_t g = 0; g < cnt; ++g) {
            uint32_t f = 0;
            #pragma unroll
            for (int p = 0; p < 8; ++p) f |= ((w[p] >> g) & 1u) << p;
            o[g] = (uint8_t)f;
        }
    }
}

// dense bytes in[(b - row0) * in_stride + g] -> plane layout (load); columns >= n become 255
__global__ void __launch_bounds__(256)
bytes_to_planes_kernel(const uint8_t* __restrict__ in, uint64_t in_stride, uint64_t row0, uint64_t nrows,
                       uint32_t n, uint8_t* __restrict__ rows, uint64_t stride) {
    const uint32_t groups = (n + 31) / 32;
    const uint64_t total = nrows * groups;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / groups;
        const uint32_t grp = (uint32_t)(i % groups);
        const uint32_t cnt = min(32u, n - grp * 32);
        const uint8_t* src = in + r * in_stride + (uint64_t)grp * 32;
        uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (uint32_t g = 0; g < 32; ++g) {
            const uint32_t f = g < cnt ? src[g] : 255u;
            #pragma unroll
            for (int p = 0; p < 8; ++p) w[p] |= ((f >> p) & 1u) << g;
        }
        *plane_ptr(rows, stride, row0 + r, grp, 0) = make_uint4(w[0], w[1], w[2], w[3]);
        *plane_ptr(rows, stride, row0 + r, grp, 1) = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

// Index merge (Miekki::merge_indexes, Miekki.cpp:901-906: the rows of the other index are appended
// column-wise): the n_src genomes of `src` (plane rows [row0, row0 + nrows) of the other index,
// staged with pitch src_stride) become columns col0 .. col0 + n_src - 1 of `rows`.  One thread per
// (row, destination group): a destination group takes its bits from two neighbouring source
// groups, funnel-shifted by col0 % 32; the first group keeps the genomes that are already there,
// columns past the last source genome stay all ones (255 = empty, Miekki.cpp:29).
__global__ void __launch_bounds__(256)
merge_planes_kernel(uint8_t* __restrict__ rows, uint64_t stride, uint32_t col0, const uint8_t* __restrict__ src,
                    uint64_t src_stride, uint32_t n_src, uint64_t row0, uint64_t nrows) {
    const uint32_t g0 = col0 >> 5, g1 = (col0 + n_src - 1) >> 5, sh = col0 & 31u;
    const uint32_t dst_groups = g1 - g0 + 1, src_groups = (n_src + 31) >> 5;
    const uint64_t total = nrows * dst_groups;
    const uint4 ones = make_uint4(~0u, ~0u, ~0u, ~0u);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / dst_groups;
        const uint32_t j = (uint32_t)(i % dst_groups);          // source group aligned with this one
        #pragma unroll
        for (int half = 0; half < 2; ++half) {
            const uint8_t* sp = src + r * src_stride + (half ? (src_stride >> 1) : 0);
            const uint4 hi = j < src_groups ? *reinterpret_cast<const uint4*>(sp + 16ull * j) : ones;
            uint4* q = plane_ptr(rows, stride, row0 + r, g0 + j, half);
            uint4 v = hi;
            if (sh) {
                uint4 lo;
                if (j) {
                    lo = *reinterpret_cast<const uint4*>(sp + 16ull * (j - 1));
                } else {
                    // genomes already in the first group: its bits below sh, moved to the top of
                    // the word the funnel shift takes them from
                    const uint4 old = *q;
                    lo = make_uint4(old.x << (32 - sh), old.y << (32 - sh), old.z << (32 - sh), old.w << (32 - sh));
                }
                v = make_uint4(__funnelshift_l(lo.x, hi.x, sh), __funnelshift_l(lo.y, hi.y, sh),
                               __funnelshift_l(lo.z, hi.z, sh), __funnelshift_l(lo.w, hi.w, sh));
            }
            *q = v;
        }
    }
}

__global__ void __launch_bounds__(256)
compact_list_kernel(const unsigned long long* __restrict__ anc, const uint8_t* __restrict__ fp,
                    SketchParams p, const uint8_t* __restrict__ bloom,
                    const uint32_t* __restrict__ read_ids, const uint64_t* __restrict__ list_off,
                    uint32_t* __restrict__ list, uint32_t* __restrict__ list_len) {
    const uint32_t s = blockIdx.y;
    const uint32_t B = 1u << p.h;
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false;
    uint32_t f = EMPTY_FP;
    if (b < B) {
        const uint64_t idx = ((uint64_t)s << p.h) + b;
        f = fp[idx];
        // Miekki.cpp:216-219: a bucket whose anc fails check_bloom is masked to 255
        keep = f != EMPTY_FP && bloom_check(bloom, p.bloom_window, p.bloom_log2, anc[idx]);
    }
    const uint32_t rid = read_ids[s];
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (m == 0) return;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(list_len + rid, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) list[list_off[rid] + base + __popc(m & ((1u << lane) - 1))] = (b << 8) | f;
}

// ---- sparse path: one CTA per read ------------------------------------------------
// shared memory: lut[256] | F words | R words | hash table (u64 slots) | counter
//   slot = bucket << 40 | fp << 32 | position ; ~0 = empty.  Linear probing; a slot is claimed
//   by its bucket with atomicCAS and then lowered with atomicMin, which orders (fp, position)
//   because the bucket bits above them are equal.
__global__ void __launch_bounds__(1024)
sketch_reads_kernel(const uint8_t* __restrict__ chars, const uint64_t* __restrict__ coff,
                    const uint64_t* __restrict__ len, const uint32_t* __restrict__ read_ids,
                    uint32_t n_ids, uint32_t max_words, uint32_t slots, SketchParams p,
                    const uint8_t* __restrict__ bloom, const uint64_t* __restrict__ list_off,
                    uint32_t* __restrict__ list, uint32_t* __restrict__ list_len) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* lut = smem;
    uint32_t* F = reinterpret_cast<uint32_t*>(smem + 256);
    uint32_t* R = F + max_words;
    unsigned long long* table =
        reinterpret_cast<unsigned long long*>(smem + 256 + ((2ull * max_words * 4 + 15) & ~15ull));
    __shared__ uint32_t cnt;
    fill_lut(lut);
    const int k = p.k, h = p.h;
    const uint64_t kmask = (1ull << (2 * k)) - 1;
    const uint64_t pmask = (1ull << (64 - h)) - 1;
    const uint32_t smask = slots - 1;

    for (uint32_t it = blockIdx.x; it < n_ids; it += gridDim.x) {
        const uint32_t rid = read_ids[it];
        const uint64_t n = len[rid];
        const uint8_t* seq = chars + coff[rid];
        __syncthreads();                          // lut ready / previous read finished
        if (threadIdx.x == 0) cnt = 0;
        const uint32_t nw = (uint32_t)((n + 15) / 16);
        const uint32_t nk = n > (uint64_t)k ? (uint32_t)(n - k) : 0;
        // table only as large as this read needs (power of two, load <= 0.75)
        uint32_t my_slots = 64;
        while ((uint64_t)my_slots * 3 < (uint64_t)nk * 4) my_slots <<= 1;
        if (my_slots > slots) my_slots = slots;
        const uint32_t my_mask = my_slots - 1;
        (void)smask;
        const bool pv = prefix_valid(seq, n, k, lut);
        for (uint32_t w = threadIdx.x; w < nw + 2; w += blockDim.x) {
            uint32_t f = 0, r = 0;
            if (w < nw) {
                const bool plain = 16ull * w >= (uint64_t)(k - 1) && 16ull * w + 16 <= n;
                if (!plain || !encode_word_acgt(*reinterpret_cast<const uint4*>(seq + 16ull * w), f, r))
                    encode_word(seq, n, w, k, pv, lut, f, r);
            }
            F[w] = f;
            R[w] = r;
        }
        for (uint32_t i = threadIdx.x; i < my_slots; i += blockDim.x) table[i] = EMPTY_KEY;
        __syncthreads();

        for (uint32_t pos = threadIdx.x; pos < nk; pos += blockDim.x) {
            const uint32_t w = pos >> 4;
            const int j = (int)(pos & 15);
            const uint64_t x = kmer_hash(F[w], F[w + 1], F[w + 2], R[w], R[w + 1], R[w + 2], j, k, kmask);
            const uint32_t fp = mantis(x & pmask, h);
            if (fp == EMPTY_FP) continue;
            const uint32_t bucket = (uint32_t)(x >> (64 - h));
            const unsigned long long key = ((unsigned long long)bucket << 40) |
                                           ((unsigned long long)fp << 32) | pos;
            uint32_t slot = (bucket * 0x9E3779B1u) >> 7 & my_mask;
            for (;;) {
                unsigned long long cur = table[slot];
                if (cur == EMPTY_KEY) {
                    cur = atomicCAS(table + slot, EMPTY_KEY, key);
                    if (cur == EMPTY_KEY) break;
                }
                if ((uint32_t)(cur >> 40) == bucket) {
                    atomicMin(table + slot, key);
                    break;
                }
                slot = (slot + 1) & my_mask;
            }
        }
        __syncthreads();

        const uint64_t out0 = list_off[rid];
        for (uint32_t i = threadIdx.x; i < my_slots; i += blockDim.x) {
            const unsigned long long e = table[i];
            if (e == EMPTY_KEY) continue;
            const uint32_t pos = (uint32_t)e;
            const uint32_t w = pos >> 4;
            const int j = (int)(pos & 15);
            const uint64_t x = kmer_hash(F[w], F[w + 1], F[w + 2], R[w], R[w + 1], R[w + 2], j, k, kmask);
            if (!bloom_check(bloom, p.bloom_window, p.bloom_log2, x)) continue;   // Miekki.cpp:217
            const uint32_t o = atomicAdd(&cnt, 1u);
            list[out0 + o] = ((uint32_t)(e >> 40) << 8) | ((uint32_t)(e >> 32) & 0xFFu);
        }
        __syncthreads();
        if (threadIdx.x == 0) list_len[rid] = cnt;
    }
}

// ---- small utilities ---------------------------------------------------------------

__global__ void fill_u64_kernel(unsigned long long* p, uint64_t n, unsigned long long v) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x)
        p[i] = v;
}
__global__ void fill_u32_kernel(uint32_t* p, uint64_t n, uint32_t v) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x)
        p[i] = v;
}

// counter-based synthetic genomes (mirrors miekki_b200/synth.py:cb_bases)
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256)
synth_kernel(uint8_t* __restrict__ chars, const uint64_t* __restrict__ coff, uint64_t seed,
             uint32_t first_g, uint64_t len) {
    const uint32_t s = blockIdx.y;
    const uint64_t base = seed + ((uint64_t)(first_g + s) << 40);
    uint8_t* dst = chars + coff[s];
    const uint64_t nq = (len + 15) / 16;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq;
         q += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t w[4] = {0, 0, 0, 0};
        #pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint64_t pos = 16 * q + i;
            const uint32_t c = pos < len ? (uint32_t)("ACGT"[splitmix64(base + pos) >> 62]) : 0u;
            w[i >> 2] |= c << (8 * (i & 3));
        }
        *reinterpret_cast<uint4*>(dst + 16 * q) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// multi-GPU Bloom merge: a byte set by a lower rank wins
__global__ void bloom_merge_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src,
                                   uint64_t n16) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 d = reinterpret_cast<uint4*>(dst)[i];
        const uint4 s = reinterpret_cast<const uint4*>(src)[i];
        uint32_t* dp = &d.x;
        const uint32_t* sp = &s.x;
        #pragma unroll
        for (int w = 0; w < 4; ++w) {
            // per byte: keep d where d != 0, else take s
            uint32_t nz = dp[w] | (dp[w] >> 4);
            nz |= nz >> 2;
            nz |= nz >> 1;
            nz &= 0x01010101u;                 // 1 per non-zero byte of d
            const uint32_t keep = nz * 0xFFu;  // 0xFF per non-zero byte
            dp[w] = (dp[w] & keep) | (sp[w] & ~keep);
        }
        reinterpret_cast<uint4*>(dst)[i] = d;
    }
}

// work counters kept on the device so that asynchronous scans need no host round trip:
// stat[0] += sum_q A(q), stat[1] += sum_q A(q) * N  (the scan's algorithmic bytes)
__global__ void __launch_bounds__(256)
account_rows_kernel(const uint32_t* __restrict__ list_len, uint32_t n, uint32_t n_genomes,
                    unsigned long long* __restrict__ stat) {
    unsigned long long s = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += list_len[i];
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) {
        atomicAdd(stat, s);
        atomicAdd(stat + 1, s * n_genomes);
    }
}

// ---- launchers ---------------------------------------------------------------------

bool smem_optin(const void* func, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> granted;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (getenv("MIEKKI_TEST_FAIL_SMEM_OPTIN")) return false;    // tests: the caller must report it, not launch
    std::lock_guard<std::mutex> lk(mu);
    size_t& have = granted[{dev, func}];
    if (bytes <= have) return true;
    if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    have = bytes;
    return true;
}

void launch_account_rows(const uint32_t* list_len, uint32_t n, uint32_t n_genomes, unsigned long long* stat,
                         cudaStream_t st) {
    if (!n) return;
    const unsigned grid = (n + 255) / 256 < 296u ? (n + 255) / 256 : 296u;
    account_rows_kernel<<<grid, 256, 0, st>>>(list_len, n, n_genomes, stat);
}

static inline unsigned blocks_for(uint64_t items, unsigned per_block, unsigned cap) {
    uint64_t b = (items + per_block - 1) / per_block;
    if (b < 1) b = 1;
    return (unsigned)(b > cap ? cap : b);
}

void launch_encode_planes(const uint8_t* chars, const uint64_t* coff, const uint64_t* len,
                          const uint64_t* woff, uint32_t n_seq, uint64_t max_len, int k,
                          uint32_t* planeF, uint32_t* planeR, cudaStream_t st) {
    if (!n_seq) return;
    const uint64_t nw = (max_len + 15) / 16 + 2;
    dim3 grid(blocks_for(nw, 256, 148 * 16), n_seq);
    encode_planes_kernel<<<grid, 256, 0, st>>>(chars, coff, len, woff, k, planeF, planeR);
}

void launch_expand_planes(const uint32_t* packed, const PackExcDev* exc, uint32_t n_exc, uint32_t seq0,
                          const uint8_t* pvalid, const uint64_t* len, const uint64_t* woff, uint32_t n_seq,
                          uint64_t max_len, int k, uint32_t* planes, cudaStream_t st) {
    if (!n_seq) return;
    const uint64_t nw = (max_len + 15) / 16 + 2;
    expand_planes_kernel<<<dim3(blocks_for(nw, 256, 148 * 16), n_seq), 256, 0, st>>>(packed, len, woff, planes);
    if (n_exc)
        patch_planes_kernel<<<(n_exc + 255) / 256, 256, 0, st>>>(exc, n_exc, seq0, len, woff, pvalid, k, planes);
}

void launch_sketch_dense(const uint32_t* planeF, const uint32_t* planeR, const uint64_t* len,
                         const uint64_t* woff, uint32_t n_seq, uint64_t max_len, int k, int h,
                         unsigned long long* keys, int ks, cudaStream_t st) {
    if (!n_seq || max_len <= (uint64_t)k) return;
    const uint64_t nwk = (max_len - k + 15) / 16;
    dim3 grid(blocks_for(nwk, 256, 148 * 16), n_seq);
    static const int prefilter = [] {
        const char* e = getenv("MIEKKI_SKETCH_PREFILTER");
        return e ? atoi(e) : 0;   // measured: no gain at -h 17, -13 % at -h 20 (a RED costs about a read)
    }();
    sketch_dense_kernel<<<grid, 256, 0, st>>>(planeF, planeR, len, woff, k, h, keys, prefilter, ks);
}

void launch_resolve(unsigned long long* keys_anc, const uint32_t* planeF, const uint32_t* planeR,
                    const uint64_t* woff, uint32_t n_seq, SketchParams p, uint8_t* fp,
                    uint32_t* active, unsigned long long* ssum, const uint8_t* bloom,
                    uint32_t* owner, const uint32_t* pair_full, uint32_t pair_words, int ks,
                    uint32_t* claims, uint32_t* n_claims, cudaStream_t st) {
    if (!n_seq) return;
    (void)planeR;
    if (pair_words > PAIR_WORDS_MAX) pair_full = nullptr;
    // up to 8 tiles per CTA, as long as the grid still fills the machine a few times over
    const uint32_t n_tiles = ((1u << p.h) + 256 * RES_U - 1) / (256 * RES_U);
    uint32_t per_cta = (uint32_t)(((uint64_t)n_tiles * n_seq) / (148u * 8u * 3u));
    per_cta = per_cta < 1 ? 1 : per_cta > 8 ? 8 : per_cta;
    dim3 grid((n_tiles + per_cta - 1) / per_cta, n_seq);
    resolve_kernel<<<grid, 256, 0, st>>>(keys_anc, planeF, woff, p, fp, active, ssum, bloom, owner, pair_full,
                                         pair_words, ks, claims, n_claims);
}

void launch_resolve_tagged(unsigned long long* keys_anc, const uint32_t* planeF, const uint64_t* woff,
                           uint32_t n_seq, SketchParams p, uint8_t* fp, uint32_t* active, unsigned long long* ssum,
                           const uint8_t* bloom, uint32_t* owner, const uint32_t* pair_full, uint32_t pair_words,
                           int ks, uint32_t* list, uint32_t* n_list, cudaStream_t st) {
    if (!n_seq) return;
    const uint32_t n_tiles = (1u << p.h) >> 10;
    uint32_t gx = (n_tiles + 1) / 2;                            // two tiles per CTA and loop trip
    const uint32_t cap = std::max(1u, 148u * 8u * 4u / n_seq);
    if (gx > cap) gx = cap;
    resolve_fast_kernel<<<dim3(gx, n_seq), 256, 0, st>>>(keys_anc, p, fp, active, ssum, pair_full, pair_words, list,
                                                        n_list);
    const uint64_t most = ((uint64_t)n_seq << p.h) / (256 * RES_U) + 1;
    resolve_slow_kernel<<<(unsigned)std::min<uint64_t>(most, 148 * 16), 256, 0, st>>>(keys_anc, planeF, woff, p, bloom,
                                                                                   owner, list, n_list, ks);
}
bool resolve_tagged_ok(int h, uint32_t pair_words) { return h >= 10 && pair_words > 0 && pair_words <= PAIR_WORDS_MAX; }

uint32_t bloom_page_count(uint64_t window) {
    return (uint32_t)((window + (1ull << BLOOM_PAGE_LOG2) - 1) >> BLOOM_PAGE_LOG2);
}

// scratch: n_pages bytes (full8) followed, 4-byte aligned, by pair words; returns the word count
uint32_t launch_bloom_pages(const uint8_t* bloom, uint64_t window, uint64_t reach, uint8_t* full8,
                            uint32_t* pair_full, cudaStream_t st) {
    const uint32_t n_pages = bloom_page_count(window);
    const uint32_t n_words = (n_pages + 31) / 32;
    if (!n_pages) return 0;
    if (reach > window) reach = window;
    bloom_pages_kernel<<<(n_pages * 32 + 255) / 256, 256, 0, st>>>(bloom, reach, n_pages, full8);
    bloom_pairs_kernel<<<(n_words + 255) / 256, 256, 0, st>>>(full8, n_pages, n_words, pair_full);
    return n_words;
}

void launch_stats_finalize(const uint32_t* active, const unsigned long long* ssum, const uint64_t* len,
                           uint32_t n, uint32_t* sketch_size, uint64_t* genome_size, float* ratio,
                           cudaStream_t st) {
    if (!n) return;
    stats_finalize_kernel<<<(n + 255) / 256, 256, 0, st>>>(active, ssum, len, n, sketch_size, genome_size, ratio);
}

void launch_bloom_commit(const unsigned long long* anc, const uint32_t* claims, const uint32_t* n_claims,
                         uint32_t n_seq, SketchParams p, uint8_t* bloom, uint32_t* owner, cudaStream_t st) {
    if (!n_seq) return;
    // the claim count lives on the device: a fixed grid strides over the list
    const uint64_t most = ((uint64_t)n_seq << p.h) / (256 * RES_U) + 1;
    bloom_commit_kernel<<<(unsigned)std::min<uint64_t>(most, 148 * 16), 256, 0, st>>>(anc, claims, n_claims, p, bloom,
                                                                                   owner);
}

void launch_scatter_planes(const uint8_t* fp, uint32_t n_seq, int h, uint8_t* rows, uint64_t stride,
                           uint32_t col0, cudaStream_t st) {
    if (!n_seq) return;
    const uint32_t groups = ((col0 + n_seq - 1) >> 5) - (col0 >> 5) + 1;
    dim3 grid((unsigned)(((2ull << h) + 255) / 256), (groups + 1) / 2);
    scatter_planes_kernel<<<grid, 256, 0, st>>>(fp, n_seq, h, rows, stride, col0);
}

void launch_planes_to_bytes(const uint8_t* rows, uint64_t stride, uint64_t row0, uint64_t nrows, uint32_t n,
                            uint8_t* out, cudaStream_t st) {
    if (!nrows || !n) return;
    const uint64_t total = nrows * ((n + 31) / 32);
    planes_to_bytes_kernel<<<blocks_for(total, 256, 148 * 32), 256, 0, st>>>(rows, stride, row0, nrows, n, out);
}

void launch_bytes_to_planes(const uint8_t* in, uint64_t in_stride, uint64_t row0, uint64_t nrows, uint32_t n,
                            uint8_t* rows, uint64_t stride, cudaStream_t st) {
    if (!nrows || !n) return;
    const uint64_t total = nrows * ((n + 31) / 32);
    bytes_to_planes_kernel<<<blocks_for(total, 256, 148 * 32), 256, 0, st>>>(in, in_stride, row0, nrows, n, rows,
                                                                          stride);
}

void launch_merge_planes(uint8_t* rows, uint64_t stride, uint32_t col0, const uint8_t* src, uint64_t src_stride,
                         uint32_t n_src, uint64_t row0, uint64_t nrows, cudaStream_t st) {
    if (!nrows || !n_src) return;
    const uint64_t total = nrows * (((col0 + n_src - 1) >> 5) - (col0 >> 5) + 1);
    merge_planes_kernel<<<blocks_for(total, 256, 148 * 32), 256, 0, st>>>(rows, stride, col0, src, src_stride, n_src,
                                                                       row0, nrows);
}

void launch_compact_list(const unsigned long long* anc, const uint8_t* fp, uint32_t n_seq,
                         SketchParams p, const uint8_t* bloom, const uint32_t* read_ids,
                         const uint64_t* list_off, uint32_t* list, uint32_t* list_len,
                         cudaStream_t st) {
    if (!n_seq) return;
    dim3 grid(((1u << p.h) + 255) / 256, n_seq);
    compact_list_kernel<<<grid, 256, 0, st>>>(anc, fp, p, bloom, read_ids, list_off, list, list_len);
}

// reads with more k-mers than this take the dense path
static constexpr uint32_t SPARSE_MAX_SLOTS = 16384;

size_t sketch_reads_smem(uint64_t max_len, int k, uint32_t* slots_out) {
    const uint64_t nk = max_len > (uint64_t)k ? max_len - k : 0;
    uint32_t slots = 64;
    while ((uint64_t)slots * 3 < nk * 4 && slots < (1u << 30)) slots <<= 1;
    if (slots_out) *slots_out = slots;
    const uint64_t words = (max_len + 15) / 16 + 2;
    return 256 + ((2 * words * 4 + 15) & ~15ull) + (size_t)slots * 8;
}

int launch_sketch_reads(const uint8_t* chars, const uint64_t* coff, const uint64_t* len,
                        const uint32_t* read_ids, uint32_t n_ids, uint64_t max_len, SketchParams p,
                        const uint8_t* bloom, const uint64_t* list_off, uint32_t* list,
                        uint32_t* list_len, cudaStream_t st) {
    if (!n_ids) return 0;
    uint32_t slots = 0;
    const size_t smem = sketch_reads_smem(max_len, p.k, &slots);
    if (slots > SPARSE_MAX_SLOTS) return -1;     // caller routes such reads to the dense path
    const uint32_t words = (uint32_t)((max_len + 15) / 16 + 2);
    if (smem > 48 * 1024 && !smem_optin(reinterpret_cast<const void*>(sketch_reads_kernel), smem)) return -2;
    const unsigned grid = n_ids < 148u * 32u ? n_ids : 148u * 32u;
    // a long read owns most of an SM's shared memory: give it enough warps to hide latency
    // (a 16,384-slot table is the only CTA on its SM)
    const unsigned threads = slots >= 16384 ? 1024u : slots >= 8192 ? 512u : slots >= 4096 ? 256u : 128u;
    sketch_reads_kernel<<<grid, threads, smem, st>>>(chars, coff, len, read_ids, n_ids, words, slots, p,
                                                 bloom, list_off, list, list_len);
    return 0;
}

void launch_fill_u64(unsigned long long* p, uint64_t n, unsigned long long v, cudaStream_t st) {
    if (!n) return;
    fill_u64_kernel<<<blocks_for(n, 256 * 8, 148 * 8), 256, 0, st>>>(p, n, v);
}
void launch_fill_u32(uint32_t* p, uint64_t n, uint32_t v, cudaStream_t st) {
    if (!n) return;
    fill_u32_kernel<<<blocks_for(n, 256 * 8, 148 * 8), 256, 0, st>>>(p, n, v);
}

void launch_synth(uint8_t* chars, const uint64_t* coff, uint64_t seed, uint32_t first_g, uint32_t n,
                  uint64_t len, cudaStream_t st) {
    if (!n || !len) return;
    dim3 grid(blocks_for((len + 15) / 16, 256, 148 * 8), n);
    synth_kernel<<<grid, 256, 0, st>>>(chars, coff, seed, first_g, len);
}

void launch_bloom_merge(uint8_t* dst, const uint8_t* src, uint64_t n, cudaStream_t st) {
    if (!n) return;
    bloom_merge_kernel<<<blocks_for(n / 16, 256 * 4, 148 * 8), 256, 0, st>>>(dst, src, n / 16);
}

}  // namespace mk
