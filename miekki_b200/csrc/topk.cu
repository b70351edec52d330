// topk.cu -- thresholds + bounded heap (Miekki::filter_results, Miekki.cpp:376-397), sm_100a.
//
// The hit list of a read is order dependent: genomes are visited in ascending id, the heap
// is libstdc++'s (push_heap / pop_heap / sort_heap, bits/stl_heap.h), a candidate replaces
// the minimum when "front > x" is false, so ties are kept or dropped by heap position
// (quirk G5).  One warp per read reproduces this exactly: 128 genomes are scored per step
// (four consecutive ids per lane), a ballot finds the lanes holding ids that may enter the
// heap, and lane 0 replays those in id order with the same sift operations.  A candidate the sequential code would skip never changes
// the heap, so skipping it early (with a possibly stale minimum, which can only be lower)
// is exact.
//
// Ties.  With saturated sketches (5 Mbp genomes at -h 17) genome_size is 0 (quirk G2), every
// intersection is 0 and, at -s 0, every genome is a candidate that replaces the heap's minimum:
// N pops and pushes per read, one at a time.  But when all K keys of a full heap equal the key of
// the candidate, every comparison of pop_heap / push_heap is false and the two calls reduce to a
// fixed move: a[p0] <- a[p1] <- ... <- a[K-1] <- candidate along the chain of right children
// p0 = 0, p1 = 2, p2 = 6, ... (tie_chain below replays __adjust_heap on equal keys once).  A step
// of 128 genomes whose m candidates all tie with such a heap is therefore applied at once: the
// chain ends up holding the last candidates of the step (shifted by m when m is shorter than the
// chain).  Any other step takes the general replay.
//
// Sharding: the heap state is read from / written to `heap`/`len`,
// so shards are chained in ascending id order; only the last link applies sort_heap.
#include "common.cuh"
#include "kernels.h"

namespace mk {

namespace {

constexpr int MAX_RESULTS = 64;
constexpr int WARPS = 4;

// comp(a, b) := a.intersection > b.intersection   (Miekki.cpp:377)

// bits/stl_heap.h:128-149 __push_heap
__device__ __forceinline__ void sift_up(HitDev* a, int hole, int top, const HitDev& v) {
    int parent = (hole - 1) / 2;
    while (hole > top && a[parent].intersection > v.intersection) {
        a[hole] = a[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a[hole] = v;
}
// bits/stl_heap.h:224-250 __adjust_heap
__device__ __forceinline__ void adjust(HitDev* a, int hole, int len, const HitDev& v) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (a[child].intersection > a[child - 1].intersection) --child;
        a[hole] = a[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a[hole] = a[child - 1];
        hole = child - 1;
    }
    sift_up(a, hole, top, v);
}
// bits/stl_heap.h:254-266 __pop_heap over a[0..n): the top moves to a[n-1]
__device__ __forceinline__ void pop(HitDev* a, int n) {
    const HitDev v = a[n - 1];
    a[n - 1] = a[0];
    adjust(a, 0, n - 1, v);
}

// positions a[p0] <- a[p1] <- ... <- a[K-1] of one pop_heap + push_heap on K equal keys
// (bits/stl_heap.h:224-250 with every comparison false); returns the chain length
__device__ int tie_chain(int K, int* pos) {
    int n = 0, child = 0;
    const int len = K - 1;
    pos[n++] = 0;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);            // the right child: "a[child] > a[child - 1]" is false
        pos[n++] = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        pos[n++] = child - 1;
    }
    if (K == 1) n = 0;                      // the popped element is the last one itself
    pos[n++] = K - 1;
    return n;
}

__global__ void __launch_bounds__(WARPS * 32)
topk_kernel(const uint32_t* __restrict__ counts, uint32_t n_reads, uint32_t n_genomes, uint32_t n_pad,
            uint32_t first_id, const uint32_t* __restrict__ sketch_size,
            const uint64_t* __restrict__ genome_size, const float* __restrict__ ratio, uint32_t K,
            uint32_t min_score, double min_intersection, HitDev* __restrict__ heap_io,
            uint32_t* __restrict__ len_io, int finalize) {
    __shared__ HitDev heaps[WARPS][MAX_RESULTS];
    __shared__ int chain[8], chain_len;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) chain_len = tie_chain((int)K, chain);
    __syncthreads();
    const uint32_t q = blockIdx.x * WARPS + warp;
    if (q >= n_reads) return;
    HitDev* hp = heaps[warp];
    uint32_t len = len_io[q];
    for (uint32_t i = lane; i < len; i += 32) hp[i] = heap_io[(uint64_t)q * K + i];
    __syncwarp();
    double hmin = len ? hp[0].intersection : 0.0;
    const int P = chain_len;
    // all K keys of the full heap are equal (to hmin): the tie shortcut applies
    auto all_equal = [&]() {
        bool eq = len >= K;
        for (uint32_t i = lane; i < len && eq; i += 32) eq = hp[i].intersection == hmin;
        return __all_sync(0xffffffffu, eq);
    };
    bool uniform = all_equal();
    const uint32_t* cq = counts + (uint64_t)q * n_pad;

    // 128 genomes per step: every lane scores four consecutive ids from one 16-byte load, and
    // the next step's load is issued before this step is evaluated (one 4-byte load per lane and
    // step left the kernel latency bound: 1 TB/s over the count tile).
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    auto load4 = [&](uint32_t g0) {
        const uint32_t base = g0 + 4u * lane;                          // n_pad is a multiple of 32
        return base < n_pad ? *reinterpret_cast<const uint4*>(cq + base) : zero4;
    };
    uint4 cur = load4(0);
    for (uint32_t g0 = 0; g0 < n_genomes; g0 += 128) {
        const uint4 nxt = g0 + 128 < n_genomes ? load4(g0 + 128) : zero4;
        const uint32_t base = g0 + 4u * lane;
        const uint32_t scv[4] = {cur.x, cur.y, cur.z, cur.w};
        double jac[4] = {0.0, 0.0, 0.0, 0.0}, x[4] = {0.0, 0.0, 0.0, 0.0};
        uint32_t mine = 0;
        #pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t g = base + j, sc = scv[j];
            if (g < n_genomes && sc >= min_score) {                    // :381
                // Cheap exact-safe screen: ratio[g] = float(genome_size / sketch_size), so
                // sc * ratio is within 3 * 2^-24 of the f64 value below.  A candidate whose
                // upper bound is still under the heap minimum would be skipped at :387 and
                // leaves no trace, so the f64 divide is spent only on possible entrants
                // (NaN / inf ratios fail the test and take the exact path).
                const double ub = (double)((float)sc * ratio[g]) * 1.000001;
                if (!(len >= K && ub < hmin)) {
                    jac[j] = (double)sc / (double)sketch_size[g];      // :382
                    x[j] = jac[j] * (double)genome_size[g];            // :383
                    // :384, and :386-387 with the heap as it stood at the start of this step
                    if (!(x[j] < min_intersection) && (len < K || !(hmin > x[j]))) mine |= 1u << j;
                }
            }
        }
        uint32_t m = __ballot_sync(0xffffffffu, mine != 0);
        if (m && uniform) {
            // every candidate of the step ties with the K equal keys of the heap?
            bool tie = true;
            #pragma unroll
            for (int j = 0; j < 4; ++j) tie = tie && (!((mine >> j) & 1u) || x[j] == hmin);
            if (__all_sync(0xffffffffu, tie)) {
                // rank of this lane's candidates among the step's, in id order
                const uint32_t cnt = __popc(mine);
                uint32_t before = cnt;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, before, o);
                    if (lane >= o) before += t;
                }
                const uint32_t total = __shfl_sync(0xffffffffu, before, 31);
                before -= cnt;
                const uint32_t used = total < (uint32_t)P ? total : (uint32_t)P;   // candidates that stay
                if (lane == 0)                                                     // shorter than the chain: shift
                    for (int i = 0; i + (int)used < P; ++i) hp[chain[i]] = hp[chain[i + used]];
                __syncwarp();
                uint32_t rank = before;
                #pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (!((mine >> j) & 1u)) continue;
                    if (rank + used >= total)                                      // one of the last `used`
                        hp[chain[P - (total - rank)]] = HitDev{first_id + base + (uint32_t)j, scv[j], jac[j], x[j]};
                    ++rank;
                }
                __syncwarp();
                m = 0;                                                             // heap stays full and uniform
            }
        }
        const bool replay = m != 0;
        while (m) {                                                    // lanes, then ids within a lane: ascending id
            const int src = __ffs((int)m) - 1;
            m &= m - 1;
            const uint32_t cm = __shfl_sync(0xffffffffu, mine, src);
            const uint32_t base_s = __shfl_sync(0xffffffffu, base, src);
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!((cm >> j) & 1u)) continue;                       // uniform over the warp
                const uint32_t sc_s = __shfl_sync(0xffffffffu, scv[j], src);
                const double j_s = __shfl_sync(0xffffffffu, jac[j], src);
                const double x_s = __shfl_sync(0xffffffffu, x[j], src);
                if (lane == 0) {
                    bool skip = false;
                    if (len >= K) {                                    // :386
                        if (hp[0].intersection > x_s) skip = true;     // :387
                        else { pop(hp, (int)len); --len; }             // :388-389
                    }
                    if (!skip) {
                        const HitDev v{first_id + base_s + (uint32_t)j, sc_s, j_s, x_s};   // :392
                        hp[len] = v;
                        ++len;
                        sift_up(hp, (int)len - 1, 0, v);               // :393
                    }
                    hmin = hp[0].intersection;
                }
                len = __shfl_sync(0xffffffffu, len, 0);
                hmin = __shfl_sync(0xffffffffu, hmin, 0);
            }
        }
        if (replay) {
            __syncwarp();
            uniform = all_equal();
        }
        cur = nxt;
    }
    if (lane == 0 && finalize)                                         // :396 sort_heap
        for (int n = (int)len; n > 1; --n) pop(hp, n);
    __syncwarp();
    for (uint32_t i = lane; i < len; i += 32) heap_io[(uint64_t)q * K + i] = hp[i];
    if (lane == 0) len_io[q] = len;
}

}  // namespace

void launch_topk(const uint32_t* counts, uint32_t n_reads, uint32_t n_genomes, uint32_t first_id,
                 const uint32_t* sketch_size, const uint64_t* genome_size, const float* ratio,
                 uint32_t nresults, uint32_t min_score, double min_intersection, HitDev* heap,
                 uint32_t* len, int finalize, cudaStream_t st) {
    if (!n_reads) return;
    const uint32_t n_pad = (n_genomes + 31) / 32 * 32;
    topk_kernel<<<(n_reads + WARPS - 1) / WARPS, WARPS * 32, 0, st>>>(
        counts, n_reads, n_genomes, n_pad, first_id, sketch_size, genome_size, ratio, nresults, min_score,
        min_intersection, heap, len, finalize);
}

}  // namespace mk
