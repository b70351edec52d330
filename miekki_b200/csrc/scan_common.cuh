// scan_common.cuh -- device helpers shared by the scan kernels (scan.cu, scan_tiled.cu):
// mbarrier / bulk-copy wrappers, the carry-save vertical counters and the bit transpose.
#pragma once
#include "common.cuh"

namespace mk {
namespace scan_detail {

constexpr int R = 4;                 // rows per carry-save block
constexpr int TOP = 16;              // counter planes: counts up to 65535 per chunk
constexpr uint32_t CHUNK_ROWS = 65504;   // rows per chunk: multiple of 32 (largest stage), < 2^16

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk async copy, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// carry-save adder: (carry, sum) of three bit-planes
__device__ __forceinline__ void csa(uint32_t& carry, uint32_t& sum, uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t u = a ^ b;
    carry = (a & b) | (u & c);
    sum = u ^ c;
}

// Vertical counters of one 32-genome group: value = ones + 2 twos + sum_l 2^l (c[l] + pend[l]),
// l = 2..TOP-1; pend[l] is occupied iff bit (l-2) of the number of R-row blocks folded so far.
// TOP_ planes hold counts below 2^TOP_ (blocks folded: at most 2^(TOP_ - 2) - 1).
template <int TOP_>
struct CountersT {
    uint32_t ones, twos;
    uint32_t c[TOP_], pend[TOP_];   // indices 0,1 unused (kept for static indexing)
    __device__ __forceinline__ void reset() {
        ones = twos = 0;
        #pragma unroll
        for (int l = 0; l < TOP_; ++l) { c[l] = 0; pend[l] = 0; }
    }
    // fold four 1-bit planes (one R-row block); nblk = blocks folded before this one
    __device__ __forceinline__ void add4(uint32_t e0, uint32_t e1, uint32_t e2, uint32_t e3, uint32_t nblk) {
        uint32_t ta, tb, carry;
        csa(ta, ones, ones, e0, e1);
        csa(tb, ones, ones, e2, e3);
        csa(carry, twos, twos, ta, tb);          // carry has weight 4
        // Binary counter with one pending slot per level: the t lowest levels whose slot is
        // occupied (t = trailing ones of nblk, the same for every thread) combine and pass a
        // carry up, the next level parks it.  Amortised one CSA per block; a jump table on t
        // keeps every case straight-line with static register indices.
#define MK_LVL(l) { uint32_t nc_; csa(nc_, c[l], c[l], pend[l], carry); carry = nc_; }
        switch (__ffs((int)~nblk) - 1) {
            case 0: pend[2] = carry; break;
            case 1: MK_LVL(2) pend[3] = carry; break;
            case 2: MK_LVL(2) MK_LVL(3) pend[4] = carry; break;
            case 3: MK_LVL(2) MK_LVL(3) MK_LVL(4) pend[5] = carry; break;
            case 4: MK_LVL(2) MK_LVL(3) MK_LVL(4) MK_LVL(5) pend[6] = carry; break;
            case 5: MK_LVL(2) MK_LVL(3) MK_LVL(4) MK_LVL(5) MK_LVL(6) pend[7] = carry; break;
            case 6: MK_LVL(2) MK_LVL(3) MK_LVL(4) MK_LVL(5) MK_LVL(6) MK_LVL(7) pend[8] = carry; break;
            case 7: MK_LVL(2) MK_LVL(3) MK_LVL(4) MK_LVL(5) MK_LVL(6) MK_LVL(7) MK_LVL(8) pend[9] = carry; break;
            case 8: MK_LVL(2) MK_LVL(3) MK_LVL(4) MK_LVL(5) MK_LVL(6) MK_LVL(7) MK_LVL(8) MK_LVL(9)
                    pend[10] = carry; break;
            case 9: MK_LVL(2) MK_LVL(3) MK_LVL(4) MK_LVL(5) MK_LVL(6) MK_LVL(7) MK_LVL(8) MK_LVL(9) MK_LVL(10)
                    pend[11] = carry; break;
            case 10: MK_LVL(2) MK_LVL(3) MK_LVL(4) MK_LVL(5) MK_LVL(6) MK_LVL(7) MK_LVL(8) MK_LVL(9) MK_LVL(10)
                     MK_LVL(11) pend[12] = carry; break;
            case 11: MK_LVL(2) MK_LVL(3) MK_LVL(4) MK_LVL(5) MK_LVL(6) MK_LVL(7) MK_LVL(8) MK_LVL(9) MK_LVL(10)
                     MK_LVL(11) MK_LVL(12) pend[13] = carry; break;
            case 12:
                if constexpr (TOP_ >= 15) {
                    MK_LVL(2) MK_LVL(3) MK_LVL(4) MK_LVL(5) MK_LVL(6) MK_LVL(7) MK_LVL(8) MK_LVL(9) MK_LVL(10)
                    MK_LVL(11) MK_LVL(12) MK_LVL(13) pend[14] = carry;
                }
                break;
            case 13:
                if constexpr (TOP_ >= 16) {
                    MK_LVL(2) MK_LVL(3) MK_LVL(4) MK_LVL(5) MK_LVL(6) MK_LVL(7) MK_LVL(8) MK_LVL(9) MK_LVL(10)
                    MK_LVL(11) MK_LVL(12) MK_LVL(13) MK_LVL(14) pend[15] = carry;
                }
                break;
            default: break;      // 14 trailing ones would need nblk >= 16383: past CHUNK_ROWS / R
        }
#undef MK_LVL
    }
    // resolve pending slots and return the TOP planes (plane l has weight 2^l)
    __device__ __forceinline__ void planes(uint32_t nblk, uint32_t (&P)[32]) {
        uint32_t carry = 0;
        P[0] = ones;
        P[1] = twos;
        #pragma unroll
        for (int l = 2; l < TOP_; ++l) {
            const uint32_t x = ((nblk >> (l - 2)) & 1u) ? pend[l] : 0u;
            uint32_t nc, s;
            csa(nc, s, c[l], x, carry);
            P[l] = s;
            carry = nc;
        }
        #pragma unroll
        for (int l = TOP_; l < 32; ++l) P[l] = 0;
    }
};
using Counters = CountersT<TOP>;

// 32 x 32 bit-matrix transpose (LSB-first): out[g] bit l = in[l] bit g
__device__ __forceinline__ void transpose32(uint32_t (&A)[32]) {
    uint32_t m = 0x0000FFFFu;
    #pragma unroll
    for (int j = 16; j != 0; j >>= 1) {
        #pragma unroll
        for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
            const uint32_t t = ((A[k] >> j) ^ A[k + j]) & m;
            A[k] ^= t << j;
            A[k + j] ^= t;
        }
        m ^= m << (j >> 1);
    }
}

}  // namespace scan_detail
}  // namespace mk
