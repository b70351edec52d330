// scan.cu -- the fingerprint scan: the HBM-bound heart of a query (sm_100a).
//
// Replaces the triple loop of Miekki::query_sequences (Miekki.cpp:355-369) and
// Miekki::query_sequence (:323-337):
//     count[q][g] = #{ (bucket, fp) of read q : rows[bucket][g] == fp }.
//
// HBM layout (see DESIGN.md "index layout"): bucket-major, one byte per (bucket, genome),
// but inside a row the eight bits of 32 consecutive genomes are stored as eight 32-bit
// bit-planes: row b = [half 0 | half 1], each half stride/2 bytes; group j (genomes
// 32j..32j+31) keeps planes 0-3 at half0 + 16j and planes 4-7 at half1 + 16j; bit i of plane p
// is bit p of the fingerprint of genome 32j+i.  One row still costs N bytes of HBM traffic,
// but "fingerprint == f" for 32 genomes is eight 3-input logic ops
//     e = AND_p (plane_p XOR m_p),   m_p = (f bit p) ? 0 : ~0,
// and the 1-bit results are summed with carry-save adders (vertical counters) instead of
// 32 byte compares + 32 adds: ~10 integer ops per 32 genome-rows.  That takes the integer
// pipe out of the way so the kernel runs at DRAM speed.
//
// Kernel: persistent, one CTA per SM.  A producer warp walks the work items (read, genome
// tile); for every R = 4 list entries it issues 2 x R bulk async copies (cp.async.bulk, SASS
// UBLKCP) of row-tile halves into one stage of a shared-memory ring guarded by mbarriers; the
// XOR masks of the R fingerprints and the item boundaries travel in a per-stage descriptor.
// Consumer threads own J groups of 32 genomes; per stage they fold R rows into the vertical
// counters; at the end of an item the planes are bit-transposed into 32 counts per group and
// written with 16-byte stores.  Algorithmic bytes = sum_q A(q) * N (one byte per compare).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "scan_common.cuh"

namespace mk {

namespace {

using namespace scan_detail;

constexpr uint32_t F_FIRST = 1, F_LAST = 2, F_END = 8, F_ACCUM = 16;
constexpr int MAX_BLOCKS = 8;        // blocks per stage: narrow tiles put up to 32 rows behind one barrier
constexpr int MAX_J = 4;
constexpr int MAX_STAGES = 48;

struct __align__(16) StageMeta {
    uint32_t flags;
    uint32_t nrows;     // rows in this stage (1..R*blocks), 0 for an empty item
    uint32_t read;
    uint32_t grp0;      // first 32-genome group of the tile
    uint32_t mask[R * MAX_BLOCKS][8];   // mask[r][p] = (fp_r bit p) ? 0 : ~0
};

template <int J>
__global__ void __launch_bounds__(384, 1)
scan_kernel(const uint8_t* __restrict__ rows, uint64_t stride, uint32_t n_groups, uint32_t n_pad,
            const uint32_t* __restrict__ list, const uint64_t* __restrict__ list_off,
            const uint32_t* __restrict__ list_len, uint32_t n_reads, uint32_t tile_groups,
            uint32_t n_tiles, int stages, uint32_t row_bytes, uint32_t half, int whole_rows, int blocks,
            uint32_t* __restrict__ counts, uint32_t* __restrict__ work_counter) {
    // stage s: R * blocks rows of row_bytes (= 32 * tile_groups: half 0 then half 1)
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t stage_rows = (uint32_t)(R * blocks);
    const uint32_t stage_bytes = stage_rows * row_bytes;
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
    // half = offset of planes 4-7 inside a stage row.  whole_rows: the tile is the whole row
    // (one genome tile, little padding), fetched with ONE bulk copy of `stride` bytes instead
    // of one per half.

    if (warp == (n_cons >> 5)) {
        // ---------------- producer warp ----------------
        int stage = 0;
        uint32_t phase = 0;
        const uint64_t n_items = (uint64_t)n_reads * n_tiles;
        for (;;) {
            unsigned long long item = 0;
            if (lane == 0) item = atomicAdd(work_counter, 1u);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= n_items) break;
            const uint32_t q = (uint32_t)(item / n_tiles);
            const uint32_t t = (uint32_t)(item % n_tiles);
            const uint32_t grp0 = t * tile_groups;
            const uint32_t gcount = (n_groups - grp0 < tile_groups) ? (n_groups - grp0) : tile_groups;
            const uint32_t hbytes = gcount * 16u;             // bytes per half row of this tile
            const uint32_t L = list_len[q];
            const uint32_t* lst = list + list_off[q];
            const uint8_t* src0 = rows + (uint64_t)grp0 * 16u;
            const uint8_t* src1 = src0 + (stride >> 1);
            uint32_t done = 0;                                // rows of this item already issued
            do {
                // lanes 0..stage_rows-1 fetch the entries of this stage (at most 32)
                const uint32_t nr = (L - done < stage_rows) ? (L - done) : stage_rows;
                const uint32_t e = ((uint32_t)lane < nr) ? __ldg(lst + done + lane) : 0u;
                const uint32_t in_chunk = done % CHUNK_ROWS;
                uint32_t flags = 0;
                if (in_chunk == 0) flags |= F_FIRST;
                if (done + nr == L || in_chunk + nr == CHUNK_ROWS) flags |= F_LAST;
                if (done >= CHUNK_ROWS) flags |= F_ACCUM;
                if (lane == 0) mbar_wait(empty + stage, phase ^ 1);
                __syncwarp();
                StageMeta* m = meta + stage;
                if ((uint32_t)lane < nr) {
                    const uint32_t fp = e & 0xFFu;
                    #pragma unroll
                    for (int p = 0; p < 8; ++p) m->mask[lane][p] = ((fp >> p) & 1u) ? 0u : 0xFFFFFFFFu;
                }
                if (lane == 0) {
                    m->flags = flags;
                    m->nrows = nr;
                    m->read = q;
                    m->grp0 = grp0;
                }
                __syncwarp();
                if (lane == 0) {
                    if (nr) mbar_arrive_expect_tx(full + stage, whole_rows ? nr * (uint32_t)stride : nr * 2u * hbytes);
                    else mbar_arrive(full + stage);
                }
                __syncwarp();
                if ((uint32_t)lane < nr) {
                    const uint64_t roff = (uint64_t)(e >> 8) * stride;
                    uint8_t* dst = ring + (size_t)stage * stage_bytes + (size_t)lane * row_bytes;
                    if (whole_rows) {
                        bulk_g2s(dst, rows + roff, (uint32_t)stride, full + stage);
                    } else {
                        bulk_g2s(dst, src0 + roff, hbytes, full + stage);
                        bulk_g2s(dst + half, src1 + roff, hbytes, full + stage);
                    }
                }
                if (++stage == stages) { stage = 0; phase ^= 1; }
                done += nr;
            } while (done < L);
        }
        if (lane == 0) {
            mbar_wait(empty + stage, phase ^ 1);
            meta[stage].flags = F_END;
            meta[stage].nrows = 0;
            mbar_arrive(full + stage);
        }
    } else {
        // ---------------- consumers ----------------
        int stage = 0;
        uint32_t phase = 0;
        Counters cnt[J];
        uint32_t nblk = 0;
        #pragma unroll
        for (int j = 0; j < J; ++j) cnt[j].reset();
        for (;;) {
            mbar_wait(full + stage, phase);
            const StageMeta* m = meta + stage;
            const uint32_t flags = m->flags;
            if (flags & F_END) break;
            const uint32_t nr = m->nrows;
            const uint32_t q = m->read, grp0 = m->grp0;
            if (flags & F_FIRST) {
                #pragma unroll
                for (int j = 0; j < J; ++j) cnt[j].reset();
                nblk = 0;
            }
            if (nr) {
                // shared-memory addresses of this thread's groups in row 0 of the stage; lanes of
                // idle groups (g >= tile_groups) read group 0 and their results are never written
                const uint32_t stage_addr = smem_u32(ring) + (uint32_t)stage * stage_bytes;
                const uint32_t mask_addr = smem_u32(&m->mask[0][0]);
                uint32_t gaddr[J];
                #pragma unroll
                for (int j = 0; j < J; ++j) {
                    const uint32_t g = (uint32_t)threadIdx.x + (uint32_t)n_cons * j;
                    gaddr[j] = stage_addr + 16u * (g < tile_groups ? g : 0u);
                }
                for (uint32_t r0 = 0; r0 < nr; r0 += R) {       // one carry-save block of R rows
                    uint32_t e[J][R];
                    const uint32_t left = nr - r0;                // >= 1; full blocks take the first branch
                    #pragma unroll
                    for (int r = 0; r < R; ++r) {
                        // branch-free: a row slot past the end of the stage re-reads the block's
                        // first row and its equality mask is zeroed below, so it contributes nothing
                        const uint32_t row = r0 + ((uint32_t)r < left ? (uint32_t)r : 0u);
                        const uint4 m0 = lds128(mask_addr + row * 32u);
                        const uint4 m1 = lds128(mask_addr + row * 32u + 16u);
                        #pragma unroll
                        for (int j = 0; j < J; ++j) {
                            const uint4 a = lds128(gaddr[j] + row * row_bytes);
                            const uint4 b = lds128(gaddr[j] + row * row_bytes + half);
                            uint32_t x = a.x ^ m0.x;
                            x = (a.y ^ m0.y) & x;
                            x = (a.z ^ m0.z) & x;
                            x = (a.w ^ m0.w) & x;
                            x = (b.x ^ m1.x) & x;
                            x = (b.y ^ m1.y) & x;
                            x = (b.z ^ m1.z) & x;
                            x = (b.w ^ m1.w) & x;
                            e[j][r] = (uint32_t)r < left ? x : 0u;
                        }
                    }
                    #pragma unroll
                    for (int j = 0; j < J; ++j) cnt[j].add4(e[j][0], e[j][1], e[j][2], e[j][3], nblk);
                    ++nblk;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + stage);       // stage may be refilled
            if (++stage == stages) { stage = 0; phase ^= 1; }
            if (flags & F_LAST) {
                const uint32_t gcount = (n_groups - grp0 < tile_groups) ? (n_groups - grp0) : tile_groups;
                #pragma unroll
                for (int j = 0; j < J; ++j) {
                    const uint32_t g = (uint32_t)threadIdx.x + (uint32_t)n_cons * j;
                    if (g < gcount) {
                        uint32_t P[32];
                        cnt[j].planes(nblk, P);
                        transpose32(P);                       // P[i] = count of genome 32 (grp0+g) + i
                        uint4* o = reinterpret_cast<uint4*>(counts + (uint64_t)q * n_pad + 32ull * (grp0 + g));
                        if (flags & F_ACCUM) {
                            #pragma unroll
                            for (int v = 0; v < 8; ++v) {
                                uint4 old = o[v];
                                o[v] = make_uint4(old.x + P[4 * v], old.y + P[4 * v + 1], old.z + P[4 * v + 2],
                                                  old.w + P[4 * v + 3]);
                            }
                        } else {
                            #pragma unroll
                            for (int v = 0; v < 8; ++v)
                                o[v] = make_uint4(P[4 * v], P[4 * v + 1], P[4 * v + 2], P[4 * v + 3]);
                        }
                    }
                }
            }
        }
    }
}

// ---- narrow shards (<= 1,024 genomes): one warp per read, no shared-memory ring -----------------
// A row of a narrow shard is a few hundred bytes: the ring kernel above then issues one bulk
// copy per row and is bound by the copy issue rate (measured 34 SM cycles per row at N = 64,
// 8.3 G rows/s on the whole GPU), not by bytes.  Here a warp owns a read; its lanes split into
// S = 32 / GP row slots x GP groups (GP = groups per row rounded up to a power of two), so one
// warp-wide 16-byte load fetches plane words of S different rows straight from L2 / HBM, four
// rows per lane in flight.  Every lane keeps carry-save counters over its own rows; at the end
// of the read the S partial counts of a group are summed with shuffles.
template <int GP>
__global__ void __launch_bounds__(256, 2)
scan_narrow_kernel(const uint8_t* __restrict__ rows, uint64_t stride, uint32_t n_groups, uint32_t n_pad,
                   const uint32_t* __restrict__ list, const uint64_t* __restrict__ list_off,
                   const uint32_t* __restrict__ list_len, uint32_t n_reads, uint32_t* __restrict__ counts,
                   uint32_t* __restrict__ work_counter) {
    constexpr int S = 32 / GP;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t g = lane & (GP - 1), rs = lane / GP;
    const bool has_group = g < n_groups;
    const uint8_t* base0 = rows + 16u * (has_group ? g : 0u);
    const uint8_t* base1 = base0 + (stride >> 1);
    for (;;) {
        uint32_t q = 0;
        if (lane == 0) q = atomicAdd(work_counter, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= n_reads) break;
        const uint32_t L = list_len[q];
        const uint32_t* lst = list + list_off[q];
        Counters cnt;
        cnt.reset();
        uint32_t nblk = 0;
        bool flushed = false;
        // counts of this lane's rows -> summed over the S row slots -> counts[q][32 g ..]
        auto flush = [&]() {
            uint32_t P[32];
            cnt.planes(nblk, P);
            transpose32(P);
            #pragma unroll
            for (int o = GP; o < 32; o <<= 1) {
                #pragma unroll
                for (int i = 0; i < 32; ++i) P[i] += __shfl_xor_sync(0xffffffffu, P[i], o);
            }
            if (rs == 0 && has_group) {
                uint4* o = reinterpret_cast<uint4*>(counts + (uint64_t)q * n_pad + 32ull * g);
                #pragma unroll
                for (int v = 0; v < 8; ++v) {
                    uint4 w = make_uint4(P[4 * v], P[4 * v + 1], P[4 * v + 2], P[4 * v + 3]);
                    if (flushed) {
                        const uint4 old = o[v];
                        w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w;
                    }
                    o[v] = w;
                }
            }
            flushed = true;
            cnt.reset();
            nblk = 0;
        };
        // entries of the next block are fetched while the rows of this one are in flight
        uint32_t en[R];
        #pragma unroll
        for (int u = 0; u < R; ++u) {
            const uint32_t idx = (uint32_t)u * S + rs;
            en[u] = idx < L ? __ldg(lst + idx) : 0xFFFFFFFFu;   // bucket < 2^24: never all ones
        }
        for (uint32_t e0 = 0; e0 < L; e0 += S * R) {
            uint32_t e[R];
            uint4 a[R], b[R];
            #pragma unroll
            for (int u = 0; u < R; ++u) {
                e[u] = en[u];
                a[u] = b[u] = make_uint4(0u, 0u, 0u, 0u);
                if (e[u] != 0xFFFFFFFFu) {
                    const uint64_t roff = (uint64_t)(e[u] >> 8) * stride;
                    a[u] = __ldg(reinterpret_cast<const uint4*>(base0 + roff));
                    b[u] = __ldg(reinterpret_cast<const uint4*>(base1 + roff));
                }
            }
            #pragma unroll
            for (int u = 0; u < R; ++u) {
                const uint32_t idx = e0 + S * R + (uint32_t)u * S + rs;
                en[u] = idx < L ? __ldg(lst + idx) : 0xFFFFFFFFu;
            }
            uint32_t x[R];
            #pragma unroll
            for (int u = 0; u < R; ++u) {
                const uint32_t fp = e[u] & 0xFFu;
                uint32_t m[8];
                #pragma unroll
                for (int p = 0; p < 8; ++p) m[p] = ((fp >> p) & 1u) ? 0u : 0xFFFFFFFFu;
                uint32_t t = a[u].x ^ m[0];
                t = (a[u].y ^ m[1]) & t;
                t = (a[u].z ^ m[2]) & t;
                t = (a[u].w ^ m[3]) & t;
                t = (b[u].x ^ m[4]) & t;
                t = (b[u].y ^ m[5]) & t;
                t = (b[u].z ^ m[6]) & t;
                t = (b[u].w ^ m[7]) & t;
                x[u] = e[u] != 0xFFFFFFFFu ? t : 0u;
            }
            cnt.add4(x[0], x[1], x[2], x[3], nblk);
            ++nblk;
            if (nblk == CHUNK_ROWS / R) flush();                // 16-plane counters are about to overflow
        }
        flush();
    }
}

}  // namespace

int scan_plan(uint32_t n_genomes, uint64_t stride, int sm_count, size_t smem_optin, int spare_sms, ScanPlan* out) {
    if (n_genomes == 0) return -1;
    const uint32_t MAXG = 512;                     // widest tile: 512 groups = 16,384 genomes
    const uint32_t G = (n_genomes + 31) / 32;      // 32-genome groups per row
    const uint32_t n_tiles = (G + MAXG - 1) / MAXG;
    const uint32_t tg = (G + n_tiles - 1) / n_tiles;
    // threads x J >= tg with the least idle lanes; prefer small J (more warps in flight)
    int best_t = 0, best_j = 0;
    uint32_t best_waste = ~0u;
    for (int j = 1; j <= MAX_J; ++j) {
        uint32_t t = (tg + j - 1) / j;
        t = (t + 31) / 32 * 32;
        if (t > 352) continue;          // 11 consumer warps + the producer warp = 384 threads
        const uint32_t waste = t * j - tg;
        if (waste < best_waste) { best_waste = waste; best_t = (int)t; best_j = j; }
    }
    if (!best_t) return -2;
    // narrow shards take the warp-per-read kernel (scan_narrow_kernel)
    // (MIEKKI_SCAN_NARROW_GROUPS=0 keeps the ring kernel; read per call so that tests can switch)
    const char* narrow_env = getenv("MIEKKI_SCAN_NARROW_GROUPS");
    const uint32_t narrow_max = (uint32_t)(narrow_env ? atoi(narrow_env) : 32);
    out->narrow_gp = 0;
    if (G <= narrow_max && G <= 32) {
        int gp = 1;
        while ((uint32_t)gp < G) gp <<= 1;
        out->narrow_gp = gp;
    }
    out->threads = best_t;
    out->J = best_j;
    out->tile_w = tg * 32;
    out->n_tiles = n_tiles;
    // one genome tile and <= 2 % padding: stage whole rows (one copy per row instead of two)
    static const bool allow_whole = [] {
        const char* e = getenv("MIEKKI_SCAN_WHOLE_ROWS");
        return !e || atoi(e) != 0;
    }();
    // (measured: +1.3 % at N = 10,000 with 1 % padding; -6 % at N = 300 where padding is 20 %)
    out->whole_rows = (allow_whole && n_tiles == 1 && (stride - (uint64_t)tg * 32) * 50 <= (uint64_t)tg * 32) ? 1 : 0;
    const size_t row_bytes = out->whole_rows ? (size_t)stride : (size_t)tg * 32;
    out->row_bytes = (uint32_t)row_bytes;
    // narrow tiles: several carry-save blocks per stage so that a stage is ~16 KB
    static const size_t stage_target = [] {
        const char* e = getenv("MIEKKI_SCAN_STAGE_TARGET_B");
        return (size_t)(e ? atoi(e) : 4096);   // measured: 16 KB stages cost resident CTAs at N = 2,000
    }();
    int blocks = (int)(stage_target / (R * row_bytes));
    blocks = std::max(1, std::min(MAX_BLOCKS, blocks));
    out->blocks = blocks;
    const size_t per_stage = (size_t)blocks * R * row_bytes + 2 * sizeof(uint64_t) + sizeof(StageMeta);
    // Narrow tiles leave an SM with two or three warps, which cannot hide their own
    // dependency chains: co-schedule several CTAs (each with its own ring and producer)
    // until about a dozen warps are resident, as long as every ring keeps >= 4 stages.
    const int warps = best_t / 32 + 1;
    int ctas = std::max(1, std::min(8, 12 / warps));
    // Leave part of every SM's shared memory to kernels that must run beside the persistent
    // scan: the top-k of the previous tile and, in sharded runs, NCCL's send/recv kernels.
    static const size_t reserve = [] {
        const char* e = getenv("MIEKKI_SCAN_SMEM_RESERVE_KB");
        return (size_t)(e ? atoi(e) : 0) * 1024;   // measured: 0 vs 40 KB -> 2 % faster step at N = 10,000
    }();
    auto ring_budget = [&](int n) { return (smem_optin - reserve - 1024) / (size_t)n - 1024; };
    while (ctas > 1 && ring_budget(ctas) / per_stage < 4) --ctas;
    int stages = (int)(ring_budget(ctas) / per_stage);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) return -3;
    out->stages = stages;
    out->smem = (size_t)stages * per_stage + 128;
    // The persistent grid fills every SM's shared memory.  A sharded run needs a few SMs for the
    // kernels that must make progress beside it (NCCL send/recv of the heap chain): they are
    // left out of the grid (mk_set_scan_spare_sms; MIEKKI_SCAN_SPARE_SMS overrides).
    if (const char* e = getenv("MIEKKI_SCAN_SPARE_SMS")) spare_sms = atoi(e);
    const int use_sms = std::max(1, sm_count - std::max(0, spare_sms));
    out->grid = use_sms * ctas;
    out->sm_count = sm_count;
    return 0;
}

template <int J>
static int launch_scan_j(const ScanPlan& plan, const uint8_t* rows, uint64_t stride,
                         uint32_t n_genomes, const uint32_t* list, const uint64_t* list_off,
                         const uint32_t* list_len, uint32_t n_reads, uint32_t* counts,
                         uint32_t* work_counter, cudaStream_t st) {
    // The opt-in is a per-device attribute of the function: a process that drives several GPUs
    // (`miekki --gpus N`, one host thread per shard) needs it on each of them.
    if (!smem_optin(reinterpret_cast<const void*>(scan_kernel<J>), plan.smem)) return -1;
    const uint32_t tg = plan.tile_w / 32;
    const uint32_t n_groups = (n_genomes + 31) / 32;
    const uint32_t n_pad = n_groups * 32;
    uint64_t items = (uint64_t)n_reads * plan.n_tiles;
    int grid = plan.grid;
    if ((uint64_t)grid > items) grid = (int)(items ? items : 1);
    scan_kernel<J><<<grid, plan.threads + 32, plan.smem, st>>>(
        rows, stride, n_groups, n_pad, list, list_off, list_len, n_reads, tg, plan.n_tiles, plan.stages,
        plan.row_bytes, plan.whole_rows ? (uint32_t)(stride / 2) : plan.row_bytes / 2, plan.whole_rows, plan.blocks,
        counts, work_counter);
    return 0;
}

template <int GP>
static int launch_narrow(const ScanPlan& plan, const uint8_t* rows, uint64_t stride, uint32_t n_genomes,
                         const uint32_t* list, const uint64_t* list_off, const uint32_t* list_len,
                         uint32_t n_reads, uint32_t* counts, uint32_t* work_counter, cudaStream_t st) {
    const uint32_t n_groups = (n_genomes + 31) / 32;
    const uint64_t warps = n_reads;                             // one read per warp and trip
    uint64_t grid = (warps + 7) / 8;
    const uint64_t cap = (uint64_t)(plan.grid > 0 ? plan.sm_count : 148) * 2;
    if (grid > cap) grid = cap;
    scan_narrow_kernel<GP><<<(unsigned)grid, 256, 0, st>>>(rows, stride, n_groups, n_groups * 32, list, list_off,
                                                           list_len, n_reads, counts, work_counter);
    return 0;
}

int launch_scan(const ScanPlan& plan, const uint8_t* rows, uint64_t stride, uint32_t n_genomes,
                const uint32_t* list, const uint64_t* list_off, const uint32_t* list_len,
                uint32_t n_reads, uint32_t* counts, uint32_t* work_counter, cudaStream_t st) {
    if (!n_reads) return 0;
    switch (plan.narrow_gp) {
        case 1: return launch_narrow<1>(plan, rows, stride, n_genomes, list, list_off, list_len, n_reads, counts, work_counter, st);
        case 2: return launch_narrow<2>(plan, rows, stride, n_genomes, list, list_off, list_len, n_reads, counts, work_counter, st);
        case 4: return launch_narrow<4>(plan, rows, stride, n_genomes, list, list_off, list_len, n_reads, counts, work_counter, st);
        case 8: return launch_narrow<8>(plan, rows, stride, n_genomes, list, list_off, list_len, n_reads, counts, work_counter, st);
        case 16: return launch_narrow<16>(plan, rows, stride, n_genomes, list, list_off, list_len, n_reads, counts, work_counter, st);
        case 32: return launch_narrow<32>(plan, rows, stride, n_genomes, list, list_off, list_len, n_reads, counts, work_counter, st);
        default: break;
    }
    switch (plan.J) {
        case 1: return launch_scan_j<1>(plan, rows, stride, n_genomes, list, list_off, list_len, n_reads, counts, work_counter, st);
        case 2: return launch_scan_j<2>(plan, rows, stride, n_genomes, list, list_off, list_len, n_reads, counts, work_counter, st);
        case 3: return launch_scan_j<3>(plan, rows, stride, n_genomes, list, list_off, list_len, n_reads, counts, work_counter, st);
        case 4: return launch_scan_j<4>(plan, rows, stride, n_genomes, list, list_off, list_len, n_reads, counts, work_counter, st);
    }
    return -1;
}

}  // namespace mk
