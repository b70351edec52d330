// scan_tiled.cu -- the fingerprint scan for read batches that cover the bucket space densely
// (long reads at small -h: BASELINE config 3, 10 kbp reads at -h 17), sm_100a.
//
// Same result as scan.cu (Miekki::query_sequences, Miekki.cpp:355-369):
//     count[q][g] = #{ (bucket, fp) of read q : rows[bucket][g] == fp }.
//
// scan.cu fetches one index row per (read, bucket) pair: N bytes of DRAM per pair, no reuse
// (ncu, round 1: DRAM bytes = 1.0004 x algorithmic, issue slots 27 % busy).  At -h 17 a 10 kbp
// read touches 7 % of the 2^17 rows, so a few dozen reads together touch nearly all of them,
// several times over: the reference itself fetches a row once per 201-read batch
// (Miekki.cpp:362).  Here a CTA owns a tile of RT = NWARPS x J = 46 reads x 1,024 genomes and
// streams ALL rows of its genome tile through shared memory once, in bucket order; every staged
// row is used by each read of the tile that has that bucket (about 3.2 of 46 at config 3).  Rows
// arrive by TMA: one 3-D tensor copy (cp.async.bulk.tensor, SASS UTMALDG) lands S = 64 rows x
// {planes 0-3, planes 4-7} x 512 B in a stage of the mbarrier ring.  CTAs take work items in
// genome-tile-major order, so the CTAs running at any time stream the same 134 MB slice of the
// index, which the 126 MB L2 serves: DRAM sees the index about once per launch instead of once
// per read (ncu: DRAM reads = 0.016 x the algorithmic bytes, L2 hit rate 93 %).
//
// Consumers: a warp owns J = 2 reads, its lanes the 32 groups (of 32 genomes) of the genome tile.
// The reads' lists are sorted by bucket (sort_lists_kernel), hold the two shared-memory offsets
// of every entry and end in sentinels; a warp keeps a 64-entry window of each list in shared
// memory and walks it in lock step with the stages: entry -> row of the stage -> two 16-byte
// loads per lane (conflict free: lane l reads bytes 16 l ..), the XOR masks of fp from a 256-entry
// table in shared memory (broadcast loads), eight 3-input logic ops, and the 1-bit results go into
// carry-save vertical counters, four rows per fold.  A fold may straddle stages: the per-read
// position inside the block of four is kept across stages (run_stage jumps back into the block).
// What bounds it: the shared-memory pipe (13 wavefronts per (read, row), 8 of them the row's 1,024
// bytes) and instruction issue; measured 14-15 TB/s of algorithmic bytes against 7-8 for scan.cu.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "kernels.h"
#include "scan_common.cuh"

namespace mk {

namespace {

using namespace scan_detail;

constexpr uint32_t T_FIRST = 1, T_LAST = 2, T_END = 8;
constexpr uint32_t ROW_BYTES = 1024;          // a staged row: planes 0-3 of 32 groups, then planes 4-7
constexpr uint32_t SENTINEL = 0xFFFFFFFFu;    // ends every sorted list: no bucket reaches 2^24
constexpr uint32_t MASK_TAB_BYTES = 256 * 32;

struct __align__(16) TileMeta {
    uint32_t flags, rt, gt, row0;
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint32_t c0, uint32_t c1, uint32_t c2,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// Vertical counters of the tiled kernel.  scan.cu's counters park one carry per level (28
// registers for 14 planes); here only the two lowest levels have a parking slot and every fourth
// block ripples its carry through the upper planes (20 logic ops per 16 rows), which brings a
// read's whole state to 23 registers -- two reads per warp at 80 registers per thread, 24 warps
// per SM.  14 planes: counts below 16,384, i.e. lists of at most TILED_MAX_ENTRIES entries.
constexpr int T_TOP = 14;
struct TileCounters {
    uint32_t ones, twos, c[T_TOP], p2, p3;      // c[0], c[1] unused
    __device__ __forceinline__ void reset() {
        ones = twos = p2 = p3 = 0;
        #pragma unroll
        for (int l = 0; l < T_TOP; ++l) c[l] = 0;
    }
    // value = ones + 2 twos + 4 (c2 + [nblk & 1] p2) + 8 (c3 + [nblk & 2] p3) + sum_{l >= 4} 2^l c[l]
    __device__ __forceinline__ void add4(uint32_t e0, uint32_t e1, uint32_t e2, uint32_t e3, uint32_t nblk) {
        uint32_t ta, tb, carry;
        csa(ta, ones, ones, e0, e1);
        csa(tb, ones, ones, e2, e3);
        csa(carry, twos, twos, ta, tb);                    // weight 4
        if ((nblk & 1u) == 0) { p2 = carry; return; }
        csa(carry, c[2], c[2], p2, carry);                 // weight 8
        if ((nblk & 2u) == 0) { p3 = carry; return; }
        csa(carry, c[3], c[3], p3, carry);                 // weight 16
        #pragma unroll
        for (int l = 4; l < T_TOP; ++l) {
            const uint32_t t = c[l] & carry;
            c[l] ^= carry;
            carry = t;
        }
    }
    __device__ __forceinline__ void planes(uint32_t nblk, uint32_t (&P)[32]) {
        uint32_t carry, nc;
        P[0] = ones;
        P[1] = twos;
        csa(carry, P[2], c[2], (nblk & 1u) ? p2 : 0u, 0u);
        csa(nc, P[3], c[3], (nblk & 2u) ? p3 : 0u, carry);
        carry = nc;
        #pragma unroll
        for (int l = 4; l < T_TOP; ++l) {
            P[l] = c[l] ^ carry;
            carry = c[l] & carry;
        }
        #pragma unroll
        for (int l = T_TOP; l < 32; ++l) P[l] = 0;
    }
};

// 14 planes overflow after 4,095 blocks of four rows: a read with a longer list (a whole genome as
// a query) has its counts written out before that (checked once per stage) and starts again from zero; the
// first write stores, later ones add (RN_ADD in ReadState::rn).  Out of line: it runs once per
// 16,380 list entries and at the end of an item.
constexpr uint32_t FLUSH_BLOCKS = 4095;
struct TileConst {                               // launch constants the flush needs, kept in shared memory
    uint32_t* counts;
    uint32_t n_pad, n_groups, n_gt, n_reads;
};
// slot = index of the read inside the tile (warp * J + j); rt / gt = the item (from the stage's meta)
__device__ __noinline__ void flush_counts(TileCounters cnt, uint32_t nblk, bool add, const TileConst* tc, uint32_t rt,
                                          uint32_t gt, uint32_t tile_reads, uint32_t slot, uint32_t lane) {
    const uint32_t q = rt * tile_reads + slot;
    const uint32_t glo = (uint32_t)(((uint64_t)gt * tc->n_groups) / tc->n_gt);
    const uint32_t ghi = (uint32_t)(((uint64_t)(gt + 1) * tc->n_groups) / tc->n_gt);
    if (q >= tc->n_reads || glo + lane >= ghi) return;      // a read past the end of the batch / a lane without a group
    uint32_t P[32];
    cnt.planes(nblk, P);
    transpose32(P);                              // P[i] = count of genome 32 (glo + lane) + i
    uint4* o = reinterpret_cast<uint4*>(tc->counts + (uint64_t)q * tc->n_pad + 32ull * (glo + lane));
    #pragma unroll
    for (int v = 0; v < 8; ++v) {
        uint4 w = make_uint4(P[4 * v], P[4 * v + 1], P[4 * v + 2], P[4 * v + 3]);
        if (add) {
            const uint4 old = o[v];
            w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w;
        }
        o[v] = w;
    }
}

// A sorted list entry is two words, the shared-memory offsets the scan needs: {bucket * ROW_BYTES,
// fp * 32}.  A warp keeps a window of 64 entries of each of its reads in shared memory (two chunks
// of 32, one entry per lane and chunk); an entry fetch is one broadcast 8-byte load, and the
// window moves on whenever the cursor enters its newer chunk (refill, out of line).
constexpr uint32_t WINDOW_BYTES = 64 * 8;

// one read of a warp: its counters, the block of four rows being gathered, its list cursor
struct ReadState {
    TileCounters cnt;
    uint32_t x0, x1, x2;          // equality planes of the rows gathered so far (r of them)
    uint2 pre;                    // this lane's entry of the chunk after the two in the window
    uint32_t ptr;                 // absolute index (in slist) of the next entry; lists start at multiples of 64
    uint32_t rn;                  // r << 24 | RN_ADD | nblk
};
constexpr uint32_t RN_ADD = 1u << 23, RN_NBLK = RN_ADD - 1;

// The window holds the chunk of `ptr` and the next one.  When ptr enters the newer chunk, the
// prefetched chunk replaces the older one and the chunk after it is requested: once per 32
// entries, out of line.
__device__ __noinline__ uint2 refill(uint2 pre, uint32_t ptr, uint32_t win, const uint2* __restrict__ slist,
                                     uint32_t lane) {
    const uint32_t cur = ptr >> 5;
    __syncwarp();                           // every lane is done with the chunk that is replaced
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(win + ((((cur + 1) & 1u) << 5 | lane) << 3)), "r"(pre.x), "r"(pre.y)
                 : "memory");
    pre = __ldg(slist + (((cur + 2) << 5) + lane));
    __syncwarp();
    return pre;
}

// next entry of the read if its bucket lies below the end of the stage
__device__ __forceinline__ bool take(ReadState& s, uint2& e, uint32_t stage_end, uint32_t win,
                                     const uint2* __restrict__ slist, uint32_t lane) {
    uint32_t lo, hi;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(win + ((s.ptr & 63u) << 3)));
    if (lo >= stage_end) return false;
    e = make_uint2(lo, hi);
    if ((++s.ptr & 31u) == 0) s.pre = refill(s.pre, s.ptr, win, slist, lane);
    return true;
}

// "fingerprint == fp" for the lane's 32 genomes of the staged row (see scan.cu for the layout)
__device__ __forceinline__ uint32_t match(uint2 e, uint32_t row_base, uint32_t mask_tab) {
    const uint32_t ra = row_base + e.x, ma = mask_tab + e.y;
    const uint4 a = lds128(ra), b = lds128(ra + 512u);
    const uint4 m0 = lds128(ma), m1 = lds128(ma + 16u);
    uint32_t x = a.x ^ m0.x;
    x = (a.y ^ m0.y) & x;
    x = (a.z ^ m0.z) & x;
    x = (a.w ^ m0.w) & x;
    x = (b.x ^ m1.x) & x;
    x = (b.y ^ m1.y) & x;
    x = (b.z ^ m1.z) & x;
    x = (b.w ^ m1.w) & x;
    return x;
}

// all entries of the read that fall into this stage; resumes inside the block of four
__device__ __forceinline__ void run_stage(ReadState& s, uint32_t stage_end, uint32_t row_base, uint32_t mask_tab,
                                          uint32_t win, const uint2* __restrict__ slist, uint32_t lane) {
    uint2 e;
    uint32_t x3;
    const uint32_t r = s.rn >> 24;
    if (r == 1) goto L1;
    if (r == 2) goto L2;
    if (r == 3) goto L3;
L0:
    if (!take(s, e, stage_end, win, slist, lane)) { s.rn &= 0xFFFFFFu; return; }   // r = 0, RN_ADD / nblk kept
    s.x0 = match(e, row_base, mask_tab);
L1:
    if (!take(s, e, stage_end, win, slist, lane)) { s.rn = (s.rn & 0xFFFFFFu) | (1u << 24); return; }
    s.x1 = match(e, row_base, mask_tab);
L2:
    if (!take(s, e, stage_end, win, slist, lane)) { s.rn = (s.rn & 0xFFFFFFu) | (2u << 24); return; }
    s.x2 = match(e, row_base, mask_tab);
L3:
    if (!take(s, e, stage_end, win, slist, lane)) { s.rn = (s.rn & 0xFFFFFFu) | (3u << 24); return; }
    x3 = match(e, row_base, mask_tab);
    s.cnt.add4(s.x0, s.x1, s.x2, x3, s.rn & RN_NBLK);
    ++s.rn;                                                       // nblk: stays below RN_ADD (flushed in time)
    goto L0;
}

// LONG: lists may exceed TILED_SHORT_ENTRIES entries; the counters are then checked once per stage
// and written out in passes (measured: 15 % slower, so reads that cannot need it do without)
template <int J, int NWARPS, bool LONG>
__global__ void __launch_bounds__((NWARPS + 1) * 32, 1)
scan_tiled_kernel(const __grid_constant__ CUtensorMap tmap, const uint2* __restrict__ slist,
                  const uint64_t* __restrict__ soff, uint32_t n_reads, uint32_t n_groups, uint32_t n_pad,
                  uint32_t n_gt, uint32_t n_rt, uint32_t n_rows, uint32_t S, int stages, uint32_t* __restrict__ counts,
                  uint32_t* __restrict__ work_counter) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t stage_bytes = S * ROW_BYTES;
    uint8_t* ring = smem;
    uint32_t* mask_tab = reinterpret_cast<uint32_t*>(smem + (size_t)stages * stage_bytes);
    uint8_t* windows = reinterpret_cast<uint8_t*>(mask_tab) + MASK_TAB_BYTES;       // NWARPS * J windows
    uint64_t* full = reinterpret_cast<uint64_t*>(windows + (size_t)NWARPS * J * WINDOW_BYTES);
    uint64_t* empty = full + stages;
    TileMeta* meta = reinterpret_cast<TileMeta*>(empty + stages);
    TileConst* tc = reinterpret_cast<TileConst*>(meta + stages);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tc->counts = counts;
        tc->n_pad = n_pad;
        tc->n_groups = n_groups;
        tc->n_gt = n_gt;
        tc->n_reads = n_reads;
    }
    // mask_tab[fp][p] = (fp bit p) ? 0 : ~0
    for (uint32_t i = threadIdx.x; i < 256 * 8; i += blockDim.x)
        mask_tab[i] = ((i >> 3) >> (i & 7)) & 1u ? 0u : 0xFFFFFFFFu;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, NWARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    constexpr uint32_t RT = NWARPS * J;           // reads per tile

    if (warp == NWARPS) {
        // ---------------- producer: one thread, one tensor copy per stage ----------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint64_t n_items = (uint64_t)n_gt * n_rt;
            for (;;) {
                const unsigned long long item = atomicAdd(work_counter, 1u);
                if (item >= n_items) break;
                // genome-tile major: CTAs that run together stream the same slice of the index
                const uint32_t gt = (uint32_t)(item / n_rt), rt = (uint32_t)(item % n_rt);
                const uint32_t glo = (uint32_t)(((uint64_t)gt * n_groups) / n_gt);
                for (uint32_t row0 = 0; row0 < n_rows; row0 += S) {
                    mbar_wait(empty + stage, phase ^ 1);
                    TileMeta* m = meta + stage;
                    m->flags = (row0 == 0 ? T_FIRST : 0u) | (row0 + S >= n_rows ? T_LAST : 0u);
                    m->rt = rt;
                    m->gt = gt;
                    m->row0 = row0;
                    mbar_arrive_expect_tx(full + stage, stage_bytes);
                    tma_load_3d(ring + (size_t)stage * stage_bytes, &tmap, glo * 4u, 0u, row0, full + stage);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
            mbar_wait(empty + stage, phase ^ 1);
            meta[stage].flags = T_END;
            mbar_arrive(full + stage);
        }
    } else {
        // ---------------- consumers: warp = J reads, lane = one group of 32 genomes ----------------
        int stage = 0;
        uint32_t phase = 0;
        ReadState rs[J];
        const uint32_t ring_addr = smem_u32(ring), mt_addr = smem_u32(mask_tab);
        const uint32_t win_addr = smem_u32(windows) + warp * J * WINDOW_BYTES;
        for (;;) {
            mbar_wait(full + stage, phase);
            const TileMeta* m = meta + stage;
            const uint32_t flags = m->flags;
            if (flags & T_END) break;
            const uint32_t row0 = m->row0, rt = m->rt, gt = m->gt;
            if (flags & T_FIRST) {
                #pragma unroll
                for (int j = 0; j < J; ++j) {
                    const uint32_t q = rt * RT + warp * J + j;
                    // reads past the end walk the sentinel block at the head of the buffer
                    const uint32_t p0 = q < n_reads ? (uint32_t)soff[q] : 0u;
                    rs[j].cnt.reset();
                    rs[j].ptr = p0;
                    rs[j].x0 = rs[j].x1 = rs[j].x2 = 0;
                    // chunks 0 and 1 into the window, chunk 2 prefetched
                    const uint2 c0 = __ldg(slist + (p0 + lane)), c1 = __ldg(slist + (p0 + 32u + lane));
                    const uint32_t w = win_addr + (uint32_t)j * WINDOW_BYTES;
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(w + (lane << 3)), "r"(c0.x), "r"(c0.y) : "memory");
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(w + ((32u + lane) << 3)), "r"(c1.x), "r"(c1.y) : "memory");
                    rs[j].pre = __ldg(slist + (p0 + 64u + lane));
                    rs[j].rn = 0;
                }
                __syncwarp();
            }
            // address of the lane's 16 bytes of row `bucket` = row_base + bucket * ROW_BYTES
            const uint32_t row_base = ring_addr + (uint32_t)stage * stage_bytes + lane * 16u - row0 * ROW_BYTES;
            const uint32_t stage_end = (row0 + S) * ROW_BYTES;
            #pragma unroll
            for (int j = 0; j < J; ++j) {
                run_stage(rs[j], stage_end, row_base, mt_addr, win_addr + (uint32_t)j * WINDOW_BYTES, slist, lane);
                // a stage adds at most S / 4 blocks: write the counts out before the planes can overflow
                if (LONG && (rs[j].rn & RN_NBLK) + S / 4 + 1 > FLUSH_BLOCKS && !(flags & T_LAST)) {
                    flush_counts(rs[j].cnt, rs[j].rn & RN_NBLK, (rs[j].rn & RN_ADD) != 0, tc, rt, gt, RT, warp * J + j, lane);
                    rs[j].cnt.reset();
                    rs[j].rn = (rs[j].rn & ~RN_NBLK) | RN_ADD;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + stage);       // stage may be refilled
            if (++stage == stages) { stage = 0; phase ^= 1; }
            if (flags & T_LAST) {
                #pragma unroll
                for (int j = 0; j < J; ++j) {
                    // the rows still waiting in the block of four (absent ones count as no match)
                    const uint32_t r = rs[j].rn >> 24;
                    uint32_t nblk = rs[j].rn & RN_NBLK;
                    if (r) {
                        rs[j].cnt.add4(rs[j].x0, r > 1 ? rs[j].x1 : 0u, r > 2 ? rs[j].x2 : 0u, 0u, nblk);
                        ++nblk;
                    }
                    if constexpr (LONG) {
                        flush_counts(rs[j].cnt, nblk, (rs[j].rn & RN_ADD) != 0, tc, rt, gt, RT, warp * J + j, lane);
                    } else {
                        const uint32_t q = rt * RT + warp * J + j;
                        const uint32_t glo = (uint32_t)(((uint64_t)gt * n_groups) / n_gt);
                        const uint32_t ghi = (uint32_t)(((uint64_t)(gt + 1) * n_groups) / n_gt);
                        if (q < n_reads && glo + lane < ghi) {
                            uint32_t P[32];
                            rs[j].cnt.planes(nblk, P);
                            transpose32(P);                   // P[i] = count of genome 32 (glo + lane) + i
                            uint4* o = reinterpret_cast<uint4*>(counts + (uint64_t)q * n_pad + 32ull * (glo + lane));
                            #pragma unroll
                            for (int v = 0; v < 8; ++v) o[v] = make_uint4(P[4 * v], P[4 * v + 1], P[4 * v + 2], P[4 * v + 3]);
                        }
                    }
                }
            }
        }
    }
}

// ---- lists sorted by bucket ------------------------------------------------------------------
// One CTA per read: the entries are scattered into shared memory -- one fingerprint byte per
// bucket plus a bitmap of the buckets that are present -- and read back in bucket order by
// walking the bitmap (2^h / 32 words, 7 % of their bits set at config 3); a run of sentinels
// follows.  Needs 2^h + 2^h / 8 bytes of shared memory: -h <= 17.  Only the bitmap is cleared.
__global__ void __launch_bounds__(1024)
sort_lists_kernel(const uint32_t* __restrict__ list, const uint64_t* __restrict__ list_off,
                  const uint32_t* __restrict__ list_len, uint32_t n_reads, uint32_t n_buckets,
                  uint2* __restrict__ slist, const uint64_t* __restrict__ soff) {
    extern __shared__ __align__(16) uint8_t dense[];                 // fp byte per bucket | bitmap
    __shared__ uint32_t warp_tot[32];
    const uint32_t padded = (n_buckets + 127) & ~127u;
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(dense + padded);
    const uint32_t n_words = padded / 32;
    // a warp owns a contiguous run of bitmap words, walked 32 words (one per lane) at a time
    const uint32_t n_warps = blockDim.x >> 5;
    const uint32_t seg = ((n_words + n_warps - 1) / n_warps + 31) & ~31u;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t w0 = min(n_words, warp * seg), w1 = min(n_words, w0 + seg);
    for (uint32_t q = blockIdx.x; q < n_reads; q += gridDim.x) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n_words; i += blockDim.x) bitmap[i] = 0;
        __syncthreads();
        const uint32_t L = list_len[q];
        const uint32_t* src = list + list_off[q];
        // four entries in flight per thread: the loads are what this loop waits for
        for (uint32_t i0 = threadIdx.x; i0 < L; i0 += 4 * blockDim.x) {
            uint32_t e[4];
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t i = i0 + u * blockDim.x;
                e[u] = i < L ? src[i] : 0xFFFFFFFFu;
            }
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i0 + u * blockDim.x >= L) continue;
                const uint32_t b = e[u] >> 8;
                dense[b] = (uint8_t)e[u];
                atomicOr(bitmap + (b >> 5), 1u << (b & 31));
            }
        }
        __syncthreads();
        uint32_t mine = 0;
        for (uint32_t w = w0 + lane; w < w1; w += 32) mine += __popc(bitmap[w]);
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if (lane == 0) warp_tot[warp] = mine;
        __syncthreads();
        // entries before this warp's run: the totals of the warps below it
        uint32_t running = lane < warp ? warp_tot[lane] : 0u;
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) running += __shfl_xor_sync(0xffffffffu, running, o);
        uint2* dst = slist + soff[q];
        for (uint32_t wb = w0; wb < w1; wb += 32) {                  // uniform trip count over the warp
            const uint32_t w = wb + lane;
            uint32_t m = w < w1 ? bitmap[w] : 0u;
            const uint32_t cnt = __popc(m);
            uint32_t incl = cnt;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (uint32_t)o) incl += t;
            }
            uint32_t at = running + incl - cnt;
            while (m) {
                const uint32_t b = 32 * w + (uint32_t)__ffs((int)m) - 1;
                m &= m - 1;
                dst[at++] = make_uint2(b * ROW_BYTES, (uint32_t)dense[b] << 5);
            }
            running += __shfl_sync(0xffffffffu, incl, 31);
        }
        // sentinels: up to the next multiple of 64 and four whole chunks more (the scan prefetches)
        const uint32_t end = ((L + 63) & ~63u) + 128;
        for (uint32_t i = L + threadIdx.x; i < end; i += blockDim.x) dst[i] = make_uint2(SENTINEL, 0u);
    }
}

__global__ void fill_sentinels_kernel(uint2* p, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = make_uint2(SENTINEL, 0u);
}

PFN_cuTensorMapEncodeTiled encode_fn() {
    static PFN_cuTensorMapEncodeTiled fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        cudaGetLastError();
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
    }();
    return fn;
}

template <int J, int NWARPS, bool LONG>
int launch_tiled_t(const TiledPlan& plan, const CUtensorMap& map, const uint2* slist, const uint64_t* soff,
                   uint32_t n_reads, uint32_t n_genomes, uint32_t n_rows, uint32_t* counts, uint32_t* work_counter,
                   cudaStream_t st) {
    if (!smem_optin(reinterpret_cast<const void*>(scan_tiled_kernel<J, NWARPS, LONG>), plan.smem)) return -1;
    const uint32_t n_groups = (n_genomes + 31) / 32;
    const uint32_t n_rt = (n_reads + plan.tile_reads - 1) / plan.tile_reads;
    const uint64_t items = (uint64_t)n_rt * plan.n_gt;
    int grid = plan.grid;
    if ((uint64_t)grid > items) grid = (int)items;
    scan_tiled_kernel<J, NWARPS, LONG><<<grid, (NWARPS + 1) * 32, plan.smem, st>>>(
        map, slist, soff, n_reads, n_groups, n_groups * 32, plan.n_gt, n_rt, n_rows, plan.S, plan.stages, counts,
        work_counter);
    return 0;
}

}  // namespace

bool tiled_scan_available() { return encode_fn() != nullptr; }

uint32_t sorted_list_capacity(uint64_t entries) { return (uint32_t)(((entries + 63) & ~63ull) + 128); }

int tiled_plan(uint32_t n_genomes, int h, int sm_count, size_t smem_optin_bytes, TiledPlan* out) {
    if (n_genomes == 0 || h > TILED_MAX_H || h < 0) return -1;
    // J reads per warp x NWARPS consumer warps: 23 registers of state per read (ReadState), 80
    // registers per thread at 768 threads
    out->J = 2;
    out->warps = 23;
    // (measured at config 3's shape: 23 / 21 / 19 consumer warps -> 161 / 168 / 169 ms per 20,000 reads)
    out->tile_reads = (uint32_t)(out->J * out->warps);
    const uint32_t n_rows = 1u << h;
    const uint32_t G = (n_genomes + 31) / 32;
    out->n_gt = (G + 31) / 32;
    // 64 rows per stage (3 stages of 64 KB): measured at 50,000 genomes, 16 / 32 / 64 rows give
    // 8.0 / 11.4 / 14.5 TB/s: the fixed cost per stage (barrier round trip, window upkeep) counts
    uint32_t S = 64;
    if (const char* e = getenv("MIEKKI_TILED_STAGE_ROWS")) S = (uint32_t)std::max(1, atoi(e));
    while (S & (S - 1)) S &= S - 1;                       // power of two
    S = std::min<uint32_t>(std::min<uint32_t>(S, 256), n_rows);
    out->S = S;
    const size_t fixed = MASK_TAB_BYTES + (size_t)out->tile_reads * WINDOW_BYTES + 256;   // 256: TileConst, alignment
    const size_t per_stage = (size_t)S * ROW_BYTES + 2 * sizeof(uint64_t) + sizeof(TileMeta);
    int stages = (int)((smem_optin_bytes - fixed) / per_stage);
    stages = std::min(stages, 64);
    if (stages < 2) return -2;
    out->stages = stages;
    out->smem = (size_t)stages * per_stage + fixed;
    out->grid = sm_count;
    return 0;
}

int launch_sort_lists(const uint32_t* list, const uint64_t* list_off, const uint32_t* list_len, uint32_t n_reads, int h,
                      void* slist_, const uint64_t* soff, cudaStream_t st) {
    uint2* slist = static_cast<uint2*>(slist_);
    if (!n_reads) return 0;
    const uint32_t n_buckets = 1u << h;
    const size_t padded = (n_buckets + 127) & ~(size_t)127;
    const size_t smem = padded + padded / 8;                       // fp bytes + bitmap
    if (smem > 48 * 1024 && !smem_optin(reinterpret_cast<const void*>(sort_lists_kernel), smem)) return -1;
    fill_sentinels_kernel<<<1, 128, 0, st>>>(slist, 128);          // the block reads past the end walk
    const unsigned grid = n_reads < 148u * 8u ? n_reads : 148u * 8u;
    // a table that leaves room for one CTA per SM only gets all the warps an SM can hold
    const unsigned threads = smem > 113 * 1024 ? 1024u : smem > 56 * 1024 ? 512u : 256u;
    sort_lists_kernel<<<grid, threads, smem, st>>>(list, list_off, list_len, n_reads, n_buckets, slist, soff);
    return 0;
}

int launch_scan_tiled(const TiledPlan& plan, const uint8_t* rows, uint64_t stride, uint32_t n_genomes, int h,
                      const void* slist_, const uint64_t* soff, uint32_t n_reads, bool long_lists, uint32_t* counts,
                      uint32_t* work_counter, cudaStream_t st) {
    if (!n_reads) return 0;
    const uint2* slist = static_cast<const uint2*>(slist_);
    PFN_cuTensorMapEncodeTiled encode = encode_fn();
    if (!encode) return -2;
    // the index as a 3-D tensor of 32-bit words: [row][half][word]; a box = S rows x 2 halves x 512 B
    CUtensorMap map;
    const uint32_t n_rows = 1u << h;
    const cuuint64_t dims[3] = {stride / 8, 2, n_rows};
    const cuuint64_t strides[2] = {stride / 2, stride};
    const cuuint32_t box[3] = {128, 2, plan.S};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t*>(rows), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return -3;
    if (long_lists)
        return launch_tiled_t<2, 23, true>(plan, map, slist, soff, n_reads, n_genomes, n_rows, counts, work_counter, st);
    return launch_tiled_t<2, 23, false>(plan, map, slist, soff, n_reads, n_genomes, n_rows, counts, work_counter, st);
}

}  // namespace mk
