// scan.cu -- the fingerprint scan: the HBM-bound heart of a query (sm_100a).
//
// Replaces the triple loop of Miekki::query_sequences (Miekki.cpp:355-369) and
// Miekki::query_sequence (:323-337):
//     count[q][g] = #{ (bucket, fp) of read q : rows[bucket][g] == fp }.
//
// Layout: rows is the bucket-major matrix, 2^h rows of `stride` bytes (stride % 128 == 0),
// one byte per genome.  A read arrives as a list of (bucket << 8 | fp) words: only buckets
// that are non-empty and passed the Bloom check.
//
// Kernel: persistent, one CTA per SM.  A producer warp walks the work items
// (read, genome tile), and for every list entry issues ONE bulk async copy
// (cp.async.bulk, SASS UBLKCP) of the row segment rows[bucket][g0 .. g0+w) into a ring of
// shared-memory stages guarded by mbarriers; its fingerprint and the item boundaries travel
// in a 16-byte per-stage descriptor.  Consumer threads own 16-byte column groups: they
// compare 4 genomes per 32-bit op (SWAR byte equality), accumulate in packed 8-bit counters
// that are spilled into 32-bit registers every 255 rows, and write the finished counters of
// an item with 16-byte stores.  Algorithmic bytes = sum_q A(q) * N (one byte per compare).
#include "common.cuh"
#include "kernels.h"

namespace mk {

namespace {

constexpr uint32_t F_FIRST = 1, F_LAST = 2, F_ROW = 4, F_END = 8;
constexpr int MAX_J = 4;
constexpr int MAX_STAGES = 64;

struct __align__(16) StageMeta {
    uint32_t splat;     // fp * 0x01010101
    uint32_t flags;
    uint32_t read;
    uint32_t g0;        // first genome of the tile
};

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

// 0x01 in every byte of x that is zero
__device__ __forceinline__ uint32_t zero_bytes(uint32_t x) {
    uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;   // bit 7 set iff low 7 bits non-zero
    t = ~(t | x) & 0x80808080u;                      // bit 7 set iff the whole byte is zero
    return t >> 7;
}

template <int J>
__global__ void __launch_bounds__(544, 1)
scan_kernel(const uint8_t* __restrict__ rows, uint64_t stride, uint32_t n_genomes, uint32_t n_pad,
            const uint32_t* __restrict__ list, const uint64_t* __restrict__ list_off,
            const uint32_t* __restrict__ list_len, uint32_t n_reads, uint32_t tile_w,
            uint32_t n_tiles, int stages, uint32_t stage_bytes, uint32_t* __restrict__ counts,
            uint32_t* __restrict__ work_counter) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* ring = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
    uint64_t* empty = full + stages;
    StageMeta* meta = reinterpret_cast<StageMeta*>(empty + stages);

    const int n_cons = blockDim.x - 32;          // consumer threads; the last warp produces
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, (uint32_t)(n_cons >> 5));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == (n_cons >> 5)) {
        // ---------------- producer warp ----------------
        int stage = 0;
        uint32_t phase = 0;
        const uint64_t n_items = (uint64_t)n_reads * n_tiles;
        auto emit = [&](uint32_t splat, uint32_t flags, uint32_t q, uint32_t g0, const uint8_t* src,
                        uint32_t bytes) {
            if (lane == 0) {
                mbar_wait(empty + stage, phase ^ 1);
                meta[stage] = StageMeta{splat, flags, q, g0};
                if (flags & F_ROW) {
                    mbar_arrive_expect_tx(full + stage, bytes);
                    bulk_g2s(ring + (size_t)stage * stage_bytes, src, bytes, full + stage);
                } else {
                    mbar_arrive(full + stage);
                }
            }
            if (++stage == stages) { stage = 0; phase ^= 1; }
        };
        for (;;) {
            unsigned long long item = 0;
            if (lane == 0) item = atomicAdd(work_counter, 1u);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= n_items) break;
            const uint32_t q = (uint32_t)(item / n_tiles);
            const uint32_t t = (uint32_t)(item % n_tiles);
            const uint32_t g0 = t * tile_w;
            const uint32_t w = (n_genomes - g0 < tile_w) ? (n_genomes - g0) : tile_w;
            const uint32_t bytes = (w + 15u) & ~15u;          // rows are padded to 128
            const uint32_t L = list_len[q];
            const uint32_t* lst = list + list_off[q];
            if (L == 0) {
                emit(0, F_FIRST | F_LAST, q, g0, nullptr, 0);
                continue;
            }
            for (uint32_t base = 0; base < L; base += 32) {
                const uint32_t mine = (base + lane < L) ? __ldg(lst + base + lane) : 0u;
                const uint32_t cnt = (L - base < 32u) ? (L - base) : 32u;
                for (uint32_t i = 0; i < cnt; ++i) {
                    const uint32_t e = __shfl_sync(0xffffffffu, mine, (int)i);
                    uint32_t flags = F_ROW;
                    if (base + i == 0) flags |= F_FIRST;
                    if (base + i == L - 1) flags |= F_LAST;
                    emit((e & 0xFFu) * 0x01010101u, flags, q, g0,
                         rows + (uint64_t)(e >> 8) * stride + g0, bytes);
                }
            }
        }
        emit(0, F_END, 0, 0, nullptr, 0);
    } else {
        // ---------------- consumers ----------------
        int stage = 0;
        uint32_t phase = 0;
        uint32_t acc8[J][4];
        uint32_t acc32[J][16];
        uint32_t pending = 0;                       // rows folded into acc8 since the last spill
        #pragma unroll
        for (int j = 0; j < J; ++j) {
            #pragma unroll
            for (int w = 0; w < 4; ++w) acc8[j][w] = 0;
            #pragma unroll
            for (int i = 0; i < 16; ++i) acc32[j][i] = 0;
        }
        auto spill = [&]() {
            #pragma unroll
            for (int j = 0; j < J; ++j) {
                #pragma unroll
                for (int w = 0; w < 4; ++w) {
                    #pragma unroll
                    for (int b = 0; b < 4; ++b) acc32[j][4 * w + b] += (acc8[j][w] >> (8 * b)) & 0xFFu;
                    acc8[j][w] = 0;
                }
            }
            pending = 0;
        };
        for (;;) {
            mbar_wait(full + stage, phase);
            const StageMeta m = meta[stage];
            if (m.flags & F_END) break;
            if (m.flags & F_FIRST) {
                #pragma unroll
                for (int j = 0; j < J; ++j) {
                    #pragma unroll
                    for (int w = 0; w < 4; ++w) acc8[j][w] = 0;
                    #pragma unroll
                    for (int i = 0; i < 16; ++i) acc32[j][i] = 0;
                }
                pending = 0;
            }
            if (m.flags & F_ROW) {
                const uint4* row = reinterpret_cast<const uint4*>(ring + (size_t)stage * stage_bytes);
                #pragma unroll
                for (int j = 0; j < J; ++j) {
                    const uint32_t col = (uint32_t)threadIdx.x + (uint32_t)n_cons * j;   // 16-byte group
                    if (col * 16u < stage_bytes) {
                        const uint4 v = row[col];
                        acc8[j][0] += zero_bytes(v.x ^ m.splat);
                        acc8[j][1] += zero_bytes(v.y ^ m.splat);
                        acc8[j][2] += zero_bytes(v.z ^ m.splat);
                        acc8[j][3] += zero_bytes(v.w ^ m.splat);
                    }
                }
                ++pending;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + stage);       // stage may be refilled
            if (++stage == stages) { stage = 0; phase ^= 1; }
            if (pending == 255u || (m.flags & F_LAST)) spill();
            if (m.flags & F_LAST) {
                const uint32_t w = (n_genomes - m.g0 < tile_w) ? (n_genomes - m.g0) : tile_w;
                uint32_t* out = counts + (uint64_t)m.read * n_pad + m.g0;
                #pragma unroll
                for (int j = 0; j < J; ++j) {
                    const uint32_t c0 = ((uint32_t)threadIdx.x + (uint32_t)n_cons * j) * 16u;
                    if (c0 < w) {                    // n_pad is a multiple of 16: whole group fits
                        uint4* o = reinterpret_cast<uint4*>(out + c0);
                        #pragma unroll
                        for (int v = 0; v < 4; ++v)
                            o[v] = make_uint4(acc32[j][4 * v], acc32[j][4 * v + 1], acc32[j][4 * v + 2],
                                              acc32[j][4 * v + 3]);
                    }
                }
            }
        }
    }
}

}  // namespace

int scan_plan(uint32_t n_genomes, int sm_count, size_t smem_optin, ScanPlan* out) {
    if (n_genomes == 0) return -1;
    const uint32_t MAXW = 16384;                   // widest tile: 512 threads x 2 groups x 16 B
    const uint32_t n16 = (n_genomes + 15) / 16;    // 16-byte groups per row
    uint32_t n_tiles = (n_genomes + MAXW - 1) / MAXW;
    uint32_t g_per_tile = (n16 + n_tiles - 1) / n_tiles;   // groups per tile
    // threads x J >= g_per_tile with the least idle lanes; prefer small J (more threads)
    int best_t = 0, best_j = 0;
    uint32_t best_waste = ~0u;
    for (int j = 1; j <= MAX_J; ++j) {
        uint32_t t = (g_per_tile + j - 1) / j;
        t = (t + 31) / 32 * 32;
        if (t < 32) t = 32;
        if (t > 512) continue;
        const uint32_t waste = t * j - g_per_tile;
        if (waste < best_waste) { best_waste = waste; best_t = (int)t; best_j = j; }
    }
    if (!best_t) return -2;
    out->threads = best_t;
    out->J = best_j;
    out->tile_w = g_per_tile * 16;
    out->n_tiles = n_tiles;
    uint32_t stage_bytes = (out->tile_w + 127) / 128 * 128;
    const size_t budget = smem_optin - 1024;       // static smem + alignment slack
    int stages = (int)(budget / (stage_bytes + 2 * sizeof(uint64_t) + sizeof(StageMeta)));
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) return -3;
    out->stages = stages;
    out->smem = (size_t)stages * (stage_bytes + 2 * sizeof(uint64_t) + sizeof(StageMeta));
    out->grid = sm_count;
    return 0;
}

template <int J>
static int launch_scan_j(const ScanPlan& plan, const uint8_t* rows, uint64_t stride,
                         uint32_t n_genomes, const uint32_t* list, const uint64_t* list_off,
                         const uint32_t* list_len, uint32_t n_reads, uint32_t* counts,
                         uint32_t* work_counter, cudaStream_t st) {
    static size_t configured = 0;
    if (plan.smem > configured) {
        if (cudaFuncSetAttribute(scan_kernel<J>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)plan.smem) != cudaSuccess)
            return -1;
        configured = plan.smem;
    }
    const uint32_t stage_bytes = (plan.tile_w + 127) / 128 * 128;
    const uint32_t n_pad = (n_genomes + 15) / 16 * 16;
    uint64_t items = (uint64_t)n_reads * plan.n_tiles;
    int grid = plan.grid;
    if ((uint64_t)grid > items) grid = (int)(items ? items : 1);
    scan_kernel<J><<<grid, plan.threads + 32, plan.smem, st>>>(
        rows, stride, n_genomes, n_pad, list, list_off, list_len, n_reads, plan.tile_w, plan.n_tiles,
        plan.stages, stage_bytes, counts, work_counter);
    return 0;
}

int launch_scan(const ScanPlan& plan, const uint8_t* rows, uint64_t stride, uint32_t n_genomes,
                const uint32_t* list, const uint64_t* list_off, const uint32_t* list_len,
                uint32_t n_reads, uint32_t* counts, uint32_t* work_counter, cudaStream_t st) {
    if (!n_reads) return 0;
    switch (plan.J) {
        case 1: return launch_scan_j<1>(plan, rows, stride, n_genomes, list, list_off, list_len, n_reads, counts, work_counter, st);
        case 2: return launch_scan_j<2>(plan, rows, stride, n_genomes, list, list_off, list_len, n_reads, counts, work_counter, st);
        case 3: return launch_scan_j<3>(plan, rows, stride, n_genomes, list, list_off, list_len, n_reads, counts, work_counter, st);
        case 4: return launch_scan_j<4>(plan, rows, stride, n_genomes, list, list_off, list_len, n_reads, counts, work_counter, st);
    }
    return -1;
}

}  // namespace mk
