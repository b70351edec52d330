// pack.cpp -- see pack.h.  Plain host C++ (no CUDA): the hot loop is an AVX2 kernel chosen at run
// time, with a portable fallback.
#include "pack.h"

#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace mk {

namespace {

// 16 characters -> packed word; false if any is not an upper-case A / C / G / T
inline bool pack16_scalar(const unsigned char* p, uint32_t& out) {
    uint32_t w = 0;
    for (int i = 0; i < 16; ++i) {
        const unsigned c = p[i];
        const unsigned x = (c >> 1) & 3u;                  // A0 C1 G3 T2
        const unsigned code = x ^ (x >> 1);                // A0 C1 G2 T3
        if ("ACGT"[code] != (char)c) return false;
        w |= code << (30 - 2 * i);
    }
    out = w;
    return true;
}

inline void add_exception(const char* s, uint64_t n, uint64_t w, std::vector<PackException>& exc) {
    PackException e;
    e.word = w;
    const uint64_t at = 16 * w;
    const uint64_t m = at < n ? (n - at < 16 ? n - at : 16) : 0;
    memcpy(e.bytes, s + at, (size_t)m);
    memset(e.bytes + m, 0, 16 - (size_t)m);
    exc.push_back(e);
}

void pack_scalar(const char* s, uint64_t w0, uint64_t w1, uint32_t* out, std::vector<PackException>& exc, uint64_t n) {
    for (uint64_t w = w0; w < w1; ++w) {
        uint32_t v = 0;
        if (!pack16_scalar(reinterpret_cast<const unsigned char*>(s) + 16 * w, v)) {
            add_exception(s, n, w, exc);
            v = 0;
        }
        out[w - w0] = v;
    }
}

#if defined(__x86_64__)
// 32 characters -> two packed words in the low 64 bits; ok = one bit per character that is A/C/G/T
__attribute__((target("avx2"), always_inline)) inline __m128i pack32_avx2(const char* p, uint32_t& ok) {
    const __m256i three = _mm256_set1_epi8(3), one = _mm256_set1_epi8(1);
    const __m256i lut = _mm256_setr_epi8('A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                         'A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i w41 = _mm256_set1_epi16(0x0104);          // bytes (4, 1): c0 * 4 + c1
    const __m256i w161 = _mm256_set1_epi32(0x00010010);     // words (16, 1): (c0 c1) * 16 + (c2 c3)
    // the low byte of dword d becomes byte 3 - d: the first four bases end up in the top byte
    const __m256i gather = _mm256_setr_epi8(12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                            12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i pair = _mm256_setr_epi32(0, 4, 0, 0, 0, 0, 0, 0);   // dwords 0 and 4 next to each other
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p));
    const __m256i x = _mm256_and_si256(_mm256_srli_epi16(v, 1), three);                 // A0 C1 G3 T2
    const __m256i code = _mm256_xor_si256(x, _mm256_and_si256(_mm256_srli_epi16(x, 1), one));
    ok = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, code), v));
    const __m256i q = _mm256_madd_epi16(_mm256_maddubs_epi16(code, w41), w161);
    return _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(_mm256_shuffle_epi8(q, gather), pair));
}

__attribute__((target("avx2"))) void pack_avx2(const char* s, uint64_t w0, uint64_t w1, uint32_t* out,
                                               std::vector<PackException>& exc, uint64_t n) {
    uint64_t w = w0;
    for (; w + 4 <= w1; w += 4) {                           // 64 characters per trip
        uint32_t ok0, ok1;
        const __m128i a = pack32_avx2(s + 16 * w, ok0), b = pack32_avx2(s + 16 * w + 32, ok1);
        _mm_storeu_si128(reinterpret_cast<__m128i*>(out + (w - w0)), _mm_unpacklo_epi64(a, b));
        if (__builtin_expect((ok0 & ok1) != 0xFFFFFFFFu, 0)) {
            const uint32_t okw[4] = {ok0 & 0xFFFFu, ok0 >> 16, ok1 & 0xFFFFu, ok1 >> 16};
            for (int i = 0; i < 4; ++i)
                if (okw[i] != 0xFFFFu) {
                    add_exception(s, n, w + i, exc);
                    out[w - w0 + i] = 0;
                }
        }
    }
    if (w < w1) pack_scalar(s, w, w1, out + (w - w0), exc, n);
}
#endif

bool have_avx2() {
#if defined(__x86_64__)
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
#else
    return false;
#endif
}

}  // namespace

const char* pack_backend() { return have_avx2() ? "avx2" : "scalar"; }

bool prefix_is_acgt(const char* s, uint64_t n, int k) {
    const uint64_t m = (uint64_t)(k - 1) < n ? (uint64_t)(k - 1) : n;
    for (uint64_t i = 0; i < m; ++i) {
        const char c = (char)(s[i] & 0xDF);                // the prefix encoder folds case
        const bool letter = (s[i] >= 'A' && s[i] <= 'Z') || (s[i] >= 'a' && s[i] <= 'z');
        if (!letter || (c != 'A' && c != 'C' && c != 'G' && c != 'T')) return false;
    }
    return true;
}

void pack_words(const char* s, uint64_t n, uint64_t w0, uint64_t w1, int k, uint32_t* out,
                std::vector<PackException>& exc) {
    // words that touch the k-1 prefix or run past the end always go to the general encoder
    const uint64_t first_plain = k > 1 ? ((uint64_t)(k - 1) + 15) / 16 : 0;
    const uint64_t end_plain = n / 16;                     // words [.., end_plain) are whole
    uint64_t w = w0;
    for (; w < w1 && w < first_plain; ++w) {
        add_exception(s, n, w, exc);
        out[w - w0] = 0;
    }
    const uint64_t mid_end = w1 < end_plain ? w1 : end_plain;
    if (w < mid_end) {
#if defined(__x86_64__)
        if (have_avx2()) pack_avx2(s, w, mid_end, out + (w - w0), exc, n);
        else
#endif
            pack_scalar(s, w, mid_end, out + (w - w0), exc, n);
        w = mid_end;
    }
    for (; w < w1; ++w) {                                  // the ragged last word (and nothing after it)
        if (16 * w < n) add_exception(s, n, w, exc);
        out[w - w0] = 0;
    }
}

}  // namespace mk
