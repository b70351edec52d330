// common.cuh -- device-side arithmetic shared by every kernel of the path.
// Each helper names the reference function whose results it must reproduce bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace mk {

constexpr uint32_t EMPTY_FP = 255u;            // Miekki.cpp:29 maximal_minimizer
constexpr uint64_t EMPTY_KEY = ~0ull;          // (fp=255, pos=max): "bucket never hit"
constexpr uint64_t EMPTY_ANC = ~0ull;          // Miekki.cpp:30 maximal_hash
constexpr int POS_BITS = 56;                   // key = fp << 56 | position << ks | tag
constexpr uint64_t POS_MASK = (1ull << POS_BITS) - 1;
// Tagged keys (sequences shorter than 2^32): ks = KEY_TAG_BITS and the low bits carry the top
// KEY_TAG_BITS bits of the canonical k-mer (the whole k-mer when 2k <= KEY_TAG_BITS).
// (KEY_TAG_BITS itself lives in kernels.h: the host picks the key format)
constexpr uint64_t KEY_TAG_MASK = (1ull << KEY_TAG_BITS) - 1;
__host__ __device__ __forceinline__ int key_tag_shift(int k) {
    return 2 * k > KEY_TAG_BITS ? 2 * k - KEY_TAG_BITS : 0;
}
// Bloom pages (sketch.cu: bloom_pages_kernel): 2^BLOOM_PAGE_LOG2 table bytes each
constexpr int BLOOM_PAGE_LOG2 = 12;

// utils.cpp:179-184 revhash64.  (x >> 32) ^ x only touches the low word, and the 64-bit
// product needs three 32-bit multiplies: lo*lo (wide), lo*hi, hi*lo.
__device__ __forceinline__ uint64_t mul_c(uint32_t lo, uint32_t hi, uint32_t clo, uint32_t chi) {
    uint64_t p = (uint64_t)lo * clo;
    uint32_t h = (uint32_t)(p >> 32) + lo * chi + hi * clo;
    return ((uint64_t)h << 32) | (uint32_t)p;
}
__device__ __forceinline__ uint64_t revhash64(uint64_t x) {
    constexpr uint32_t CLO = 0x6659FD93u, CHI = 0xD6E8FEB8u;
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    x = mul_c(lo ^ hi, hi, CLO, CHI);
    lo = (uint32_t)x; hi = (uint32_t)(x >> 32);
    x = mul_c(lo ^ hi, hi, CLO, CHI);
    lo = (uint32_t)x; hi = (uint32_t)(x >> 32);
    return ((uint64_t)hi << 32) | (lo ^ hi);
}
// utils.cpp:188-193 unrevhash64 (the inverse permutation)
__device__ __forceinline__ uint64_t unrevhash64(uint64_t x) {
    constexpr uint32_t CLO = 0x8B59A89Bu, CHI = 0xCFEE444Du;
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    x = mul_c(lo ^ hi, hi, CLO, CHI);
    lo = (uint32_t)x; hi = (uint32_t)(x >> 32);
    x = mul_c(lo ^ hi, hi, CLO, CHI);
    lo = (uint32_t)x; hi = (uint32_t)(x >> 32);
    return ((uint64_t)hi << 32) | (lo ^ hi);
}

// Miekki.cpp:91-113 mantis for number_bit_minimizer = 8, number_bit_mantis = 5:
// fp = max(p - 32 + h, 0) << 3 | (3 bits below the leading one), p = floor(log2 n); 0 -> 255.
__device__ __forceinline__ uint32_t mantis(uint64_t n, int h) {
    if (n == 0) return EMPTY_FP;
    const int lz = __clzll((long long)n);
    const int p = 63 - lz;
    int e = p - 32 + h;
    e = e < 0 ? 0 : e;
    uint32_t suf;
    if (p >= 3) suf = (uint32_t)((n << lz) >> 60) & 7u;   // (n - 2^p) >> (p - 3)
    else suf = (uint32_t)n - (1u << p);                    // offset clamps at 0
    return ((uint32_t)e << 3) + suf;                       // e <= 31 because p <= 63 - h
}

// Miekki.cpp:121-146 + utils.cpp:197-199: Bloom slot of probe i for the bucket hash x.
// universal_hash(x, i) = unrevhash64(x) + ((uint32)(i*69) * revhash64(x)) % 1024
struct BloomProbe {
    uint64_t base;      // unrevhash64(x): the canonical k-mer itself
    uint32_t rlow;      // revhash64(x) mod 1024 is all the i-term needs
    __device__ __forceinline__ explicit BloomProbe(uint64_t x)
        : base(unrevhash64(x)), rlow((uint32_t)revhash64(x) & 1023u) {}
    __device__ __forceinline__ uint64_t slot(uint32_t i, uint32_t b) const {
        return (base + (uint64_t)((i * 69u * rlow) & 1023u)) >> b;
    }
};

// The probe offsets are < 1024 < 2^(b+3), so the five probes fall in at most two adjacent
// table bytes.  For claiming / writing a byte only the FIRST probe that maps to it matters
// (smallest probe index wins, Miekki.cpp:122-129 in probe order).  Returns the number of
// distinct bytes (1 or 2); slot[j] / probe[j] describe that first probe.
__device__ __forceinline__ int bloom_first_probes(const BloomProbe& pr, uint32_t b, uint64_t (&slot)[2],
                                                  uint32_t (&probe)[2]) {
    slot[0] = pr.slot(0, b);
    probe[0] = 0;
    int n = 1;
    #pragma unroll
    for (uint32_t i = 1; i < 5; ++i) {
        const uint64_t s = pr.slot(i, b);
        if (n == 1 && (s >> 3) != (slot[0] >> 3)) {
            slot[1] = s;
            probe[1] = i;
            n = 2;
        }
    }
    return n;
}

// Bloom_Filter bytes are only ever zero or one power of two (insert writes 1 << bit into
// a zero byte, Miekki.cpp:127-128); membership is "all five bytes non-zero" (:141).
__device__ __forceinline__ bool bloom_check(const uint8_t* __restrict__ table, uint64_t window,
                                            uint32_t b, uint64_t x) {
    BloomProbe pr(x);
    uint64_t last = ~0ull;
    #pragma unroll
    for (uint32_t i = 0; i < 5; ++i) {
        uint64_t byte = pr.slot(i, b) >> 3;
        if (byte == last) continue;                 // the five probes almost always coincide
        if (byte >= window || table[byte] == 0) return false;
        last = byte;
    }
    return true;
}

// ---- packed sequence planes ------------------------------------------------------
// Every sequence is re-encoded once into two 2-bit planes of 32-bit words (16 bases each):
//   F (forward digits, big-endian: base j at bits 30-2*(j%16)) and
//   R (reverse-strand digits, little-endian: base j at bits 2*(j%16)),
// so that the forward k-mer starting at base i is a left-aligned bit field of F and the
// reverse-complement k-mer a right-aligned bit field of R.  The digits are exactly what the
// reference's rolling update would have shifted in: nuc2int / nuc2intrc (utils.cpp:31-49,
// 107-125) for bases at index >= k-1, and the case-insensitive, all-or-nothing prefix
// encoder str2numstrand (utils.cpp:252-272) with rcb (Miekki.cpp:66-76) below k-1.

// 64-bit window of 32 bases starting at base (16*w + j), big-endian, from three F words
__device__ __forceinline__ uint64_t fwd_window(uint32_t w0, uint32_t w1, uint32_t w2, int j) {
    const uint32_t hi = __funnelshift_l(w1, w0, 2 * j);
    const uint32_t lo = __funnelshift_l(w2, w1, 2 * j);
    return ((uint64_t)hi << 32) | lo;
}
// 64-bit window of 32 bases starting at base (16*w + j), little-endian, from three R words
__device__ __forceinline__ uint64_t rev_window(uint32_t r0, uint32_t r1, uint32_t r2, int j) {
    const uint32_t lo = __funnelshift_r(r0, r1, 2 * j);
    const uint32_t hi = __funnelshift_r(r1, r2, 2 * j);
    return ((uint64_t)hi << 32) | lo;
}

// canonical k-mer (Miekki.cpp:167) of the k-mer starting at base (16*w + j), and its hash (:168)
__device__ __forceinline__ uint64_t kmer_canon(uint32_t f0, uint32_t f1, uint32_t f2,
                                               uint32_t r0, uint32_t r1, uint32_t r2,
                                               int j, int k, uint64_t kmask) {
    const uint64_t S = fwd_window(f0, f1, f2, j) >> (64 - 2 * k);
    const uint64_t RC = rev_window(r0, r1, r2, j) & kmask;
    return S < RC ? S : RC;
}
__device__ __forceinline__ uint64_t kmer_hash(uint32_t f0, uint32_t f1, uint32_t f2,
                                              uint32_t r0, uint32_t r1, uint32_t r2,
                                              int j, int k, uint64_t kmask) {
    return revhash64(kmer_canon(f0, f1, f2, r0, r1, r2, j, k, kmask));
}

}  // namespace mk
