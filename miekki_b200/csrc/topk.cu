// topk.cu -- thresholds + bounded heap (Miekki::filter_results, Miekki.cpp:376-397), sm_100a.
//
// The hit list of a read is order dependent: genomes are visited in ascending id, the heap
// is libstdc++'s (push_heap / pop_heap / sort_heap, bits/stl_heap.h), a candidate replaces
// the minimum when "front > x" is false, so ties are kept or dropped by heap position
// (quirk G5).  One warp per read reproduces this exactly: 32 genomes are scored per step,
// a ballot finds the ones that may enter the heap, and lane 0 replays those in id order
// with the same sift operations.  A candidate the sequential code would skip never changes
// the heap, so skipping it early (with a possibly stale minimum, which can only be lower)
// is exact.
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

__global__ void __launch_bounds__(WARPS * 32)
topk_kernel(const uint32_t* __restrict__ counts, uint32_t n_reads, uint32_t n_genomes, uint32_t n_pad,
            uint32_t first_id, const uint32_t* __restrict__ sketch_size,
            const uint64_t* __restrict__ genome_size, const float* __restrict__ ratio, uint32_t K,
            uint32_t min_score, double min_intersection, HitDev* __restrict__ heap_io,
            uint32_t* __restrict__ len_io, int finalize) {
    __shared__ HitDev heaps[WARPS][MAX_RESULTS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x * WARPS + warp;
    if (q >= n_reads) return;
    HitDev* hp = heaps[warp];
    uint32_t len = len_io[q];
    for (uint32_t i = lane; i < len; i += 32) hp[i] = heap_io[(uint64_t)q * K + i];
    __syncwarp();
    double hmin = len ? hp[0].intersection : 0.0;
    const uint32_t* cq = counts + (uint64_t)q * n_pad;

    for (uint32_t g0 = 0; g0 < n_genomes; g0 += 32) {
        const uint32_t g = g0 + lane;
        bool cand = false;
        uint32_t sc = 0;
        double jac = 0.0, x = 0.0;
        if (g < n_genomes) {
            sc = cq[g];
            if (sc >= min_score) {                                     // :381
                // Cheap exact-safe screen: ratio[g] = float(genome_size / sketch_size), so
                // sc * ratio is within 3 * 2^-24 of the f64 value below.  A candidate whose
                // upper bound is still under the heap minimum would be skipped at :387 and
                // leaves no trace, so the f64 divide is spent only on possible entrants
                // (NaN / inf ratios fail the test and take the exact path).
                const double ub = (double)((float)sc * ratio[g]) * 1.000001;
                if (!(len >= K && ub < hmin)) {
                    jac = (double)sc / (double)sketch_size[g];         // :382
                    x = jac * (double)genome_size[g];                  // :383
                    cand = !(x < min_intersection);                    // :384
                }
            }
        }
        // :386-387 with the heap as it stood at the start of this step
        const bool maybe = cand && (len < K || !(hmin > x));
        uint32_t m = __ballot_sync(0xffffffffu, maybe);
        while (m) {
            const int src = __ffs((int)m) - 1;
            m &= m - 1;
            const uint32_t g_s = __shfl_sync(0xffffffffu, g, src);
            const uint32_t sc_s = __shfl_sync(0xffffffffu, sc, src);
            const double j_s = __shfl_sync(0xffffffffu, jac, src);
            const double x_s = __shfl_sync(0xffffffffu, x, src);
            if (lane == 0) {
                bool skip = false;
                if (len >= K) {                                        // :386
                    if (hp[0].intersection > x_s) skip = true;         // :387
                    else { pop(hp, (int)len); --len; }                 // :388-389
                }
                if (!skip) {
                    const HitDev v{first_id + g_s, sc_s, j_s, x_s};    // :392
                    hp[len] = v;
                    ++len;
                    sift_up(hp, (int)len - 1, 0, v);                   // :393
                }
                hmin = hp[0].intersection;
            }
            len = __shfl_sync(0xffffffffu, len, 0);
            hmin = __shfl_sync(0xffffffffu, hmin, 0);
        }
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
