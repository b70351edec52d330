// exact.cu -- exact mode (-e): true k-mer intersection of candidate hits, sm_100a.
//
// Replaces Miekki::ground_truth_batch (Miekki.cpp:792-859): set B = distinct canonical
// k-mers of the genome's records, and per read A = its distinct canonical k-mers,
// nb_inter = |A n B|, nb_union = |B| + |A \ B|.  Canonical k-mer = str2num
// (utils.cpp:276-278): min(str2numstrand(w), str2numstrand(revComp(w))), case-insensitive,
// and 0 for any window holding a byte outside ACGTacgt (str2numstrand returns 0, revComp
// maps the byte to 'T', so the minimum is 0).  Every window 0 .. len-k is included
// (unlike the sketch loop, quirk G1).
//
// The reference uses std::unordered_set; here both sets are exact open-addressing hash
// sets of 64-bit keys in HBM/L2 (insert-if-absent with atomicCAS, empty = ~0, which is not
// a k-mer for k <= 31).  |B| is the number of successful inserts.
#include <cstdlib>

#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "kernels.h"

namespace mk {

namespace {

constexpr unsigned long long EMPTY = ~0ull;
// Windows per thread in the genome kernel.  Every insert is a chain of two dependent random
// accesses (probe, CAS) that only other threads can hide: ncu at 64 windows per thread showed
// 24 % occupancy (306 CTAs for 5 Mbp), 14 % issue utilisation, 250 us per 5 Mbp genome.  Short
// runs re-roll k - 1 characters more often but fill the machine.
static int run_length() {
    static const int v = [] {
        const char* e = getenv("MIEKKI_EXACT_RUN");
        const int r = e ? atoi(e) : 8;     // measured, 5 Mbp genome: 64 -> 250 us, 8 -> 158 us
        return r < 1 ? 1 : r;
    }();
    return v;
}

// digit (0..3) of str2numstrand, 4 for a byte it rejects (utils.cpp:252-272)
__device__ __forceinline__ uint32_t ci_code(uint32_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
    }
    return 4;
}

// insert-if-absent; true when the key was new
__device__ __forceinline__ bool set_insert(unsigned long long* table, uint64_t slots, uint64_t v) {
    uint64_t i = __umul64hi(revhash64(v ^ 0x5851F42D4C957F2Dull), slots);
    for (;;) {
        unsigned long long cur = table[i];
        if (cur == v) return false;
        if (cur == EMPTY) {
            cur = atomicCAS(table + i, EMPTY, (unsigned long long)v);
            if (cur == EMPTY) return true;
            if (cur == v) return false;
        }
        if (++i == slots) i = 0;
    }
}
__device__ __forceinline__ bool set_contains(const unsigned long long* __restrict__ table,
                                             uint64_t slots, uint64_t v) {
    uint64_t i = __umul64hi(revhash64(v ^ 0x5851F42D4C957F2Dull), slots);
    for (;;) {
        const unsigned long long cur = table[i];
        if (cur == v) return true;
        if (cur == EMPTY) return false;
        if (++i == slots) i = 0;
    }
}

// genome records -> set B.  One thread rolls over RUN consecutive windows.
__global__ void __launch_bounds__(256)
exact_insert_kernel(const uint8_t* __restrict__ chars, const uint64_t* __restrict__ coff,
                    const uint64_t* __restrict__ len, int k, unsigned long long* __restrict__ table,
                    uint64_t slots, unsigned long long* __restrict__ distinct, int RUN) {
    const uint32_t s = blockIdx.y;
    const uint64_t n = len[s];
    uint32_t fresh = 0;
    if (n >= (uint64_t)k) {
        const uint64_t nwin = n - k + 1;                       // Miekki.cpp:807
        const uint8_t* seq = chars + coff[s];
        const uint64_t kmask = (1ull << (2 * k)) - 1;
        for (uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * RUN; i0 < nwin;
             i0 += (uint64_t)gridDim.x * blockDim.x * RUN) {
            const uint64_t iend = (i0 + RUN < nwin) ? i0 + RUN : nwin;
            uint64_t fwd = 0, rc = 0;
            int64_t last_bad = -1;
            for (uint64_t t = i0; t < iend + k - 1; ++t) {
                uint32_t c = ci_code(seq[t]);
                if (c > 3) { last_bad = (int64_t)t; c = 0; }
                fwd = ((fwd << 2) | c) & kmask;
                rc = (rc >> 2) | ((uint64_t)(3 - c) << (2 * k - 2));
                if (t + 1 >= i0 + k) {
                    const uint64_t i = t + 1 - k;
                    const uint64_t v = (last_bad >= (int64_t)i) ? 0ull : (fwd < rc ? fwd : rc);
                    fresh += set_insert(table, slots, v) ? 1u : 0u;
                }
            }
        }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) fresh += __shfl_xor_sync(0xffffffffu, fresh, o);
    if ((threadIdx.x & 31) == 0 && fresh) atomicAdd(distinct, (unsigned long long)fresh);
}

// reads: thread per window; per-read set A lives in rtable[toff[r] .. toff[r] + tslots[r])
__global__ void __launch_bounds__(256)
exact_reads_kernel(const uint8_t* __restrict__ chars, const uint64_t* __restrict__ coff,
                   const uint64_t* __restrict__ len, int k, unsigned long long* __restrict__ rtable,
                   const uint64_t* __restrict__ toff, const unsigned long long* __restrict__ tableB,
                   uint64_t slotsB, unsigned long long* __restrict__ inter,
                   unsigned long long* __restrict__ distinctA) {
    const uint32_t r = blockIdx.y;
    const uint64_t n = len[r];
    uint32_t n_new = 0, n_in = 0;
    if (n >= (uint64_t)k) {
        const uint64_t nwin = n - k + 1;                       // Miekki.cpp:832
        const uint8_t* seq = chars + coff[r];
        unsigned long long* setA = rtable + toff[r];
        const uint64_t slotsA = toff[r + 1] - toff[r];
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwin;
             i += (uint64_t)gridDim.x * blockDim.x) {
            uint64_t fwd = 0, rc = 0;
            bool bad = false;
            for (int t = 0; t < k; ++t) {
                uint32_t c = ci_code(seq[i + t]);
                if (c > 3) { bad = true; c = 0; }
                fwd = (fwd << 2) | c;
                rc = (rc >> 2) | ((uint64_t)(3 - c) << (2 * k - 2));
            }
            const uint64_t v = bad ? 0ull : (fwd < rc ? fwd : rc);
            if (set_insert(setA, slotsA, v)) {                 // :834 first time in A
                ++n_new;
                if (set_contains(tableB, slotsB, v)) ++n_in;   // :835
            }
        }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_new += __shfl_xor_sync(0xffffffffu, n_new, o);
        n_in += __shfl_xor_sync(0xffffffffu, n_in, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_new) atomicAdd(distinctA + r, (unsigned long long)n_new);
        if (n_in) atomicAdd(inter + r, (unsigned long long)n_in);
    }
}

// ---- the sort-merge variant (BASELINE.json north_star (3) words exact mode as "a sorted-hash
// merge"; the reference has an unused precedent, main.cpp:82-98 collisions_sort) ----------------
// Set B as a sorted array: every window's canonical k-mer is written out, radix-sorted (the one
// library call of this file: cub::DeviceRadixSort over the 2k significant bits), |B| = number of
// positions that differ from their predecessor, membership = binary search.  Kept beside the
// hash sets for the comparison DESIGN.md section 4 reports; MIEKKI_EXACT_SORT=1 selects it.
__global__ void __launch_bounds__(256)
exact_keys_kernel(const uint8_t* __restrict__ chars, const uint64_t* __restrict__ coff,
                  const uint64_t* __restrict__ len, const uint64_t* __restrict__ koff, int k,
                  unsigned long long* __restrict__ keys, int RUN) {
    const uint32_t s = blockIdx.y;
    const uint64_t n = len[s];
    if (n < (uint64_t)k) return;
    const uint64_t nwin = n - k + 1;
    const uint8_t* seq = chars + coff[s];
    unsigned long long* out = keys + koff[s];
    const uint64_t kmask = (1ull << (2 * k)) - 1;
    for (uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * RUN; i0 < nwin;
         i0 += (uint64_t)gridDim.x * blockDim.x * RUN) {
        const uint64_t iend = (i0 + RUN < nwin) ? i0 + RUN : nwin;
        uint64_t fwd = 0, rc = 0;
        int64_t last_bad = -1;
        for (uint64_t t = i0; t < iend + k - 1; ++t) {
            uint32_t c = ci_code(seq[t]);
            if (c > 3) { last_bad = (int64_t)t; c = 0; }
            fwd = ((fwd << 2) | c) & kmask;
            rc = (rc >> 2) | ((uint64_t)(3 - c) << (2 * k - 2));
            if (t + 1 >= i0 + k) {
                const uint64_t i = t + 1 - k;
                out[i] = (last_bad >= (int64_t)i) ? 0ull : (fwd < rc ? fwd : rc);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
exact_count_distinct_kernel(const unsigned long long* __restrict__ sorted, uint64_t n,
                            unsigned long long* __restrict__ distinct) {
    uint32_t mine = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        mine += (i == 0 || sorted[i] != sorted[i - 1]) ? 1u : 0u;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(distinct, (unsigned long long)mine);
}

__device__ __forceinline__ bool sorted_contains(const unsigned long long* __restrict__ a, uint64_t n, uint64_t v) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return lo < n && a[lo] == v;
}

// exact_reads_kernel with set B as a sorted array
__global__ void __launch_bounds__(256)
exact_reads_sorted_kernel(const uint8_t* __restrict__ chars, const uint64_t* __restrict__ coff,
                          const uint64_t* __restrict__ len, int k, unsigned long long* __restrict__ rtable,
                          const uint64_t* __restrict__ toff, const unsigned long long* __restrict__ sortedB,
                          uint64_t nB, unsigned long long* __restrict__ inter,
                          unsigned long long* __restrict__ distinctA) {
    const uint32_t r = blockIdx.y;
    const uint64_t n = len[r];
    uint32_t n_new = 0, n_in = 0;
    if (n >= (uint64_t)k) {
        const uint64_t nwin = n - k + 1;
        const uint8_t* seq = chars + coff[r];
        unsigned long long* setA = rtable + toff[r];
        const uint64_t slotsA = toff[r + 1] - toff[r];
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwin;
             i += (uint64_t)gridDim.x * blockDim.x) {
            uint64_t fwd = 0, rc = 0;
            bool bad = false;
            for (int t = 0; t < k; ++t) {
                uint32_t c = ci_code(seq[i + t]);
                if (c > 3) { bad = true; c = 0; }
                fwd = (fwd << 2) | c;
                rc = (rc >> 2) | ((uint64_t)(3 - c) << (2 * k - 2));
            }
            const uint64_t v = bad ? 0ull : (fwd < rc ? fwd : rc);
            if (set_insert(setA, slotsA, v)) {
                ++n_new;
                if (sorted_contains(sortedB, nB, v)) ++n_in;
            }
        }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_new += __shfl_xor_sync(0xffffffffu, n_new, o);
        n_in += __shfl_xor_sync(0xffffffffu, n_in, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_new) atomicAdd(distinctA + r, (unsigned long long)n_new);
        if (n_in) atomicAdd(inter + r, (unsigned long long)n_in);
    }
}

}  // namespace

size_t exact_sort_temp_bytes(uint64_t n_keys, int k) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, static_cast<const unsigned long long*>(nullptr),
                                   static_cast<unsigned long long*>(nullptr), n_keys, 0, 2 * k);
    return bytes;
}

void launch_exact_sorted_build(const uint8_t* chars, const uint64_t* coff, const uint64_t* len, const uint64_t* koff,
                               uint32_t n_seq, uint64_t max_len, uint64_t n_keys, int k, unsigned long long* keys,
                               unsigned long long* sorted, void* temp, size_t temp_bytes, unsigned long long* distinct,
                               cudaStream_t st) {
    if (!n_seq || !n_keys) return;
    const int RUN = run_length();
    const uint64_t threads = (max_len - k + 1 + RUN - 1) / RUN;
    uint64_t bx = (threads + 255) / 256;
    if (bx > 148 * 16) bx = 148 * 16;
    exact_keys_kernel<<<dim3((unsigned)bx, n_seq), 256, 0, st>>>(chars, coff, len, koff, k, keys, RUN);
    cub::DeviceRadixSort::SortKeys(temp, temp_bytes, keys, sorted, n_keys, 0, 2 * k, st);
    uint64_t cb = (n_keys + 255) / 256;
    if (cb > 148 * 8) cb = 148 * 8;
    exact_count_distinct_kernel<<<(unsigned)cb, 256, 0, st>>>(sorted, n_keys, distinct);
}

void launch_exact_reads_sorted(const uint8_t* chars, const uint64_t* coff, const uint64_t* len, uint32_t n_reads,
                               uint64_t max_len, int k, unsigned long long* rtable, const uint64_t* toff,
                               const unsigned long long* sortedB, uint64_t nB, unsigned long long* inter,
                               unsigned long long* distinctA, cudaStream_t st) {
    if (!n_reads || max_len < (uint64_t)k) return;
    uint64_t bx = (max_len - k + 1 + 255) / 256;
    if (bx > 148 * 4) bx = 148 * 4;
    exact_reads_sorted_kernel<<<dim3((unsigned)bx, n_reads), 256, 0, st>>>(chars, coff, len, k, rtable, toff, sortedB,
                                                                          nB, inter, distinctA);
}

void launch_exact_insert(const uint8_t* chars, const uint64_t* coff, const uint64_t* len,
                         uint32_t n_seq, uint64_t max_len, int k, unsigned long long* table,
                         uint64_t slots, unsigned long long* distinct, cudaStream_t st) {
    if (!n_seq || max_len < (uint64_t)k) return;
    const int RUN = run_length();
    const uint64_t threads = (max_len - k + 1 + RUN - 1) / RUN;
    uint64_t bx = (threads + 255) / 256;
    if (bx > 148 * 16) bx = 148 * 16;
    dim3 grid((unsigned)bx, n_seq);
    exact_insert_kernel<<<grid, 256, 0, st>>>(chars, coff, len, k, table, slots, distinct, RUN);
}

void launch_exact_reads(const uint8_t* chars, const uint64_t* coff, const uint64_t* len,
                        uint32_t n_reads, uint64_t max_len, int k, unsigned long long* rtable,
                        const uint64_t* toff, const unsigned long long* tableB, uint64_t slotsB,
                        unsigned long long* inter, unsigned long long* distinctA, cudaStream_t st) {
    if (!n_reads || max_len < (uint64_t)k) return;
    uint64_t bx = (max_len - k + 1 + 255) / 256;
    if (bx > 148 * 4) bx = 148 * 4;
    dim3 grid((unsigned)bx, n_reads);
    exact_reads_kernel<<<grid, 256, 0, st>>>(chars, coff, len, k, rtable, toff, tableB, slotsB, inter,
                                             distinctA);
}

}  // namespace mk
