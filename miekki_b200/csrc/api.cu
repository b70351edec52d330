// api.cu -- the C ABI (include/miekki_b200.h): context, HBM layout, orchestration.
//
// HBM layout of one context (one GPU, one genome shard):
//   rows        2^h x stride bytes, bucket-major, stride % 128 == 0, columns >= n hold 255
//               (== vector<string> index, Miekki.h:54, with the genome axis contiguous)
//   sketch_size u32[cap], genome_size u64[cap]                       (Miekki.h:59-60)
//   bloom       the first `window` bytes of the 2^b/8-byte table      (Miekki.h:56)
//   owner       u32[window]: scratch for the order-exact Bloom insert (all ~0 between calls)
// plus grow-only scratch for sequence planes, bucket keys, query lists and count tiles.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <future>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/miekki_b200.h"
#include "kernels.h"
#include "pack.h"

using namespace mk;

static_assert(sizeof(mk_hit) == 24 && sizeof(HitDev) == 24, "mk_hit layout");

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

enum Phase { PH_SKETCH = 0, PH_READ_SKETCH, PH_SCAN, PH_TOPK, PH_EXACT, PH_COUNT };

struct PendingEvent {
    cudaEvent_t a, b;
    int phase;
};

}  // namespace

struct mk_batch {
    uint8_t* chars = nullptr;     // device; sequence i at chars + h_coff[i] (16-byte aligned)
    uint64_t* d_coff = nullptr;   // device copies of the two arrays below (same allocation)
    uint64_t* d_len = nullptr;
    cudaStream_t stream = nullptr;   // stream the allocation is ordered on
    std::vector<uint64_t> h_coff, h_len;
    uint32_t n = 0;
    uint64_t bytes = 0;           // padded size of chars
    uint64_t bases = 0, max_len = 0;
    uint64_t serial = next_serial();   // tells a batch from a later one at the same address
    static uint64_t next_serial() {
        static std::atomic<uint64_t> n{0};
        return ++n;
    }
};

struct mk_ctx {
    uint32_t k, h, nbm, nbmant, b, threshold;
    bool key_tags = true;         // MIEKKI_KEY_TAGS=0 forces the plain key format (used for >= 2^32-base sequences)
    uint64_t B;
    int device = 0, sm_count = 148;
    int scan_spare_sms = 0;       // SMs the persistent scan leaves to concurrent kernels (mk_set_scan_spare_sms)
    size_t smem_optin = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t aux_stream = nullptr;     // top-k of one count tile overlaps the scan of the next
    cudaStream_t sk_stream = nullptr;      // read sketch of the next tile overlaps the scan too
    cudaStream_t bl_stream = nullptr;      // build_lists: stream override (nullptr = main stream)
    int bl_set = 0;                        // build_lists: which list buffer set to fill
    std::mutex mu;
    std::string err;

    uint8_t* rows = nullptr;
    uint64_t stride = 0;
    uint32_t n = 0, cap = 0, first_id = 0;
    uint32_t importing = 0;       // mk_index_import_begin .. _end: genomes of the index being loaded
    bool import_open = false;
    uint32_t* d_sketch_size = nullptr;
    uint64_t* d_genome_size = nullptr;
    float* d_ratio = nullptr;     // float(genome_size / sketch_size), screen of the top-k kernel
    // host mirrors of the two arrays above, valid for ids [0, h_sketch_size.size()) and
    // completed from the device on demand (host_stats): the build never waits for them
    std::vector<uint32_t> h_sketch_size;
    std::vector<uint64_t> h_genome_size;
    uint8_t* bloom = nullptr;
    uint32_t* owner = nullptr;
    uint64_t window = 0;          // bytes, multiple of 16
    uint64_t bloom_reach = 0;     // bytes [bloom_reach, window) can never be probed (mk_bloom_reach)

    DevBuf planeF, planeR, keys, fp, meta, list, list_len, list2, list_len2, counts, counts2, heap, heap_len, heap2,
        heap_len2, misc, pages, slist, slist2;
    void* pinned = nullptr;
    size_t pinned_cap = 0;
    uint32_t* d_work = nullptr;
    // mk_scan_async / mk_topk_slot: two count tiles (counts, counts2) and heap scratches
    int next_slot = 0, last_slot = 0;
    uint32_t slot_reads[2] = {0, 0};
    cudaEvent_t slot_ev[2] = {nullptr, nullptr};
    bool slot_used[2] = {false, false};
    cudaEvent_t sk_ev = nullptr;
    // mk_sketch_async: lists of the batch the next mk_scan_async will be given, built ahead of it
    struct PreparedLists {
        uint64_t serial = 0;
        int slot = -1;
        uint32_t *list = nullptr, *list_len = nullptr;
        uint64_t *list_off = nullptr, *soff = nullptr;
        void* slist = nullptr;
        bool long_lists = false;
        cudaEvent_t ev = nullptr;
    } prep;
    void* meta_pin[2] = {nullptr, nullptr};
    size_t meta_pin_cap[2] = {0, 0};
    int meta_flip = 0;
    cudaStream_t copy_stream = nullptr;                              // uploads that overlap compute
    std::atomic<uint64_t> h2d_meta{0};                               // bytes counted off the main thread
    cudaEvent_t meta_ev[2] = {nullptr, nullptr};
    unsigned long long* d_stat = nullptr;   // device work counters: rows, row bytes

    std::vector<cudaEvent_t> ev_pool;
    std::vector<PendingEvent> ev_pending;
    mk_stats stats{};

    SketchParams sp() const { return SketchParams{(int)k, (int)h, b, window}; }
};

namespace {

int fail(mk_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    else g_create_error = msg;
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(c, e_ == cudaErrorMemoryAllocation ? MK_ERR_NOMEM : MK_ERR_CUDA,      \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                  \
    } while (0)

#define TRY(expr)                  \
    do {                           \
        int r_ = (expr);           \
        if (r_ != MK_OK) return r_; \
    } while (0)

int reserve(mk_ctx* c, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return MK_OK;
    if (b.p) {
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 4 + 256;
    if (cudaMalloc(&b.p, want) != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        CU(cudaMalloc(&b.p, want));
    }
    b.cap = want;
    return MK_OK;
}

int reserve_pinned(mk_ctx* c, size_t bytes) {
    if (bytes <= c->pinned_cap) return MK_OK;
    if (c->pinned) {
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaFreeHost(c->pinned));
        c->pinned = nullptr;
        c->pinned_cap = 0;
    }
    CU(cudaMallocHost(&c->pinned, bytes + bytes / 4));
    c->pinned_cap = bytes + bytes / 4;
    return MK_OK;
}

// ---- device timing ---------------------------------------------------------------
cudaEvent_t get_event(mk_ctx* c) {
    if (!c->ev_pool.empty()) {
        cudaEvent_t e = c->ev_pool.back();
        c->ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
struct PhaseTimer {
    mk_ctx* c;
    PendingEvent pe;
    cudaStream_t st;
    PhaseTimer(mk_ctx* ctx, int phase, cudaStream_t stream = nullptr) : c(ctx), st(stream ? stream : ctx->stream) {
        pe.a = get_event(c);
        pe.b = get_event(c);
        pe.phase = phase;
        cudaEventRecord(pe.a, st);
    }
    ~PhaseTimer() {
        cudaEventRecord(pe.b, st);
        c->ev_pending.push_back(pe);
    }
};
void resolve_events(mk_ctx* c) {
    for (auto& pe : c->ev_pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, pe.a, pe.b) == cudaSuccess) {
            double* slot[PH_COUNT] = {&c->stats.sketch_ms, &c->stats.read_sketch_ms, &c->stats.scan_ms,
                                      &c->stats.topk_ms, &c->stats.exact_ms};
            *slot[pe.phase] += ms;
        }
        c->ev_pool.push_back(pe.a);
        c->ev_pool.push_back(pe.b);
    }
    c->ev_pending.clear();
}
int sync(mk_ctx* c) {
    CU(cudaStreamSynchronize(c->stream));
    resolve_events(c);
    return MK_OK;
}

// ---- index storage ---------------------------------------------------------------
int ensure_capacity(mk_ctx* c, uint32_t need) {
    if (need <= c->cap) return MK_OK;
    uint32_t ncap = std::max<uint32_t>(need, c->cap ? c->cap * 2 : 128);
    ncap = (ncap + 127) / 128 * 128;
    uint8_t* nrows = nullptr;
    const uint64_t bytes = c->B * (uint64_t)ncap;
    if (cudaMalloc(&nrows, bytes) != cudaSuccess) {
        cudaGetLastError();
        ncap = (need + 127) / 128 * 128;       // retry without the growth slack
        CU(cudaMalloc(&nrows, c->B * (uint64_t)ncap));
    }
    CU(cudaMemsetAsync(nrows, 0xFF, c->B * (uint64_t)ncap, c->stream));
    uint32_t* nss = nullptr;
    uint64_t* ngs = nullptr;
    CU(cudaMalloc(&nss, (size_t)ncap * 4));
    CU(cudaMalloc(&ngs, (size_t)ncap * 8));
    CU(cudaMemsetAsync(nss, 0, (size_t)ncap * 4, c->stream));
    CU(cudaMemsetAsync(ngs, 0, (size_t)ncap * 8, c->stream));
    if (c->n) {
        // a row is [half 0 | half 1]: each half moves to its place in the wider row
        const size_t used = (size_t)((c->n + 31) / 32) * 16;     // bytes in use per half
        CU(cudaMemcpy2DAsync(nrows, ncap, c->rows, c->stride, used, c->B, cudaMemcpyDeviceToDevice,
                             c->stream));
        CU(cudaMemcpy2DAsync(nrows + ncap / 2, ncap, c->rows + c->stride / 2, c->stride, used, c->B,
                             cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemcpyAsync(nss, c->d_sketch_size, (size_t)c->n * 4, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemcpyAsync(ngs, c->d_genome_size, (size_t)c->n * 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    float* nratio = nullptr;
    CU(cudaMalloc(&nratio, (size_t)ncap * 4));
    CU(cudaMemsetAsync(nratio, 0, (size_t)ncap * 4, c->stream));
    if (c->n) CU(cudaMemcpyAsync(nratio, c->d_ratio, (size_t)c->n * 4, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (c->rows) CU(cudaFree(c->rows));
    if (c->d_sketch_size) CU(cudaFree(c->d_sketch_size));
    if (c->d_genome_size) CU(cudaFree(c->d_genome_size));
    if (c->d_ratio) CU(cudaFree(c->d_ratio));
    c->rows = nrows;
    c->d_sketch_size = nss;
    c->d_genome_size = ngs;
    c->d_ratio = nratio;
    c->cap = ncap;
    c->stride = ncap;
    return MK_OK;
}

// ratio[g] = float(genome_size / sketch_size) for ids [first, first + n): the top-k screen
int upload_ratio(mk_ctx* c, uint32_t first, uint32_t n) {
    if (!n) return MK_OK;
    std::vector<float> r(n);
    for (uint32_t i = 0; i < n; ++i)
        r[i] = (float)((double)c->h_genome_size[first + i] / (double)c->h_sketch_size[first + i]);
    CU(cudaMemcpyAsync(c->d_ratio + first, r.data(), (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->stats.h2d_bytes += (size_t)n * 4;
    return MK_OK;
}

// ---- batches ---------------------------------------------------------------------
void batch_layout(mk_batch* b, const uint64_t* lens, uint32_t n) {
    b->n = n;
    b->h_coff.resize((size_t)n + 1);
    b->h_len.assign(lens, lens + n);
    uint64_t off = 0;
    b->bases = 0;
    b->max_len = 0;
    for (uint32_t i = 0; i < n; ++i) {
        b->h_coff[i] = off;
        off += (lens[i] + 15) / 16 * 16;
        b->bases += lens[i];
        b->max_len = std::max(b->max_len, lens[i]);
    }
    b->h_coff[n] = off;
    b->bytes = off + 64;          // kernels may read one 16-byte word past a sequence
}

// Per-call buffers come from the stream-ordered pool (release threshold raised in
// mk_create): cudaMalloc/cudaFree per query batch cost up to hundreds of ms on a GPU
// that holds a multi-GB index.
int batch_alloc(mk_ctx* c, mk_batch* b, cudaStream_t st = nullptr) {
    if (!st) st = c->stream;
    b->stream = st;
    // one allocation: chars | coff[n+1] | len[n]
    const size_t chars_bytes = (b->bytes + 255) / 256 * 256;
    const size_t total = chars_bytes + ((size_t)b->n + 1) * 8 + std::max<size_t>(1, b->n) * 8;
    CU(cudaMallocAsync(reinterpret_cast<void**>(&b->chars), total, st));
    b->d_coff = reinterpret_cast<uint64_t*>(b->chars + chars_bytes);
    b->d_len = b->d_coff + b->n + 1;
    CU(cudaMemcpyAsync(b->d_coff, b->h_coff.data(), ((size_t)b->n + 1) * 8, cudaMemcpyHostToDevice,
                       st));
    if (b->n)
        CU(cudaMemcpyAsync(b->d_len, b->h_len.data(), (size_t)b->n * 8, cudaMemcpyHostToDevice, st));
    c->h2d_meta.fetch_add(((size_t)b->n * 2 + 1) * 8, std::memory_order_relaxed);
    return MK_OK;
}

void batch_release(mk_batch* b) {
    if (!b) return;
    if (b->chars) cudaFreeAsync(b->chars, b->stream);
    delete b;
}

// gathers sequences [first, first+n) of (seqs, lens) into a new device batch
// `st`: stream to copy on (default: the ctx stream).  With an explicit stream the function may
// run on a helper thread next to the main one (mk_index_add overlaps the upload of the next
// genomes with the sketching of the current ones): it then touches only its own staging ring,
// ring events and atomic counters.
int upload_range(mk_ctx* c, const char* const* seqs, const uint64_t* lens, uint32_t n, mk_batch** out,
                 cudaStream_t st = nullptr) {
    if (!st) st = c->stream;
    mk_batch* b = new mk_batch();
    batch_layout(b, lens, n);
    int r = batch_alloc(c, b, st);
    if (r != MK_OK) { batch_release(b); return r; }
    // zero the alignment gaps once so that padding bytes are deterministic
    cudaError_t e = cudaMemsetAsync(b->chars, 0, b->bytes, st);
    if (e != cudaSuccess) { batch_release(b); return fail(c, MK_ERR_CUDA, cudaGetErrorString(e)); }
    const bool big = n && (b->bases / n) >= (1u << 20);
    if (big) {
        // Few long sequences (genomes) in pageable memory.  One host thread copies ~14 GB/s into
        // pinned memory and PCIe 5 x16 takes 55 GB/s (measured on this pool), so a small team
        // stages pieces of different sequences concurrently: every worker owns two pinned pieces
        // (copy into one while the other is in flight over PCIe) and issues its own async copies.
        static const size_t PIECE = [] {
            const char* e = getenv("MIEKKI_UPLOAD_PIECE_MB");
            return (size_t)(e ? std::max(1, atoi(e)) : 8) << 20;
        }();
        static const unsigned TEAM = [] {
            const char* e = getenv("MIEKKI_UPLOAD_THREADS");
            const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
            return (unsigned)std::min(16, std::max(1, e ? atoi(e) : (int)std::min(6u, std::max(1u, hw / 2))));
        }();
        struct Piece { uint32_t seq; uint64_t off; size_t bytes; };
        std::vector<Piece> pieces;
        for (uint32_t i = 0; i < n; ++i)
            for (uint64_t off = 0; off < lens[i]; off += PIECE)
                pieces.push_back({i, off, (size_t)std::min<uint64_t>(PIECE, lens[i] - off)});
        const unsigned team = (unsigned)std::min<size_t>(TEAM, std::max<size_t>(1, pieces.size()));
        r = reserve_pinned(c, PIECE * 2 * team);
        if (r != MK_OK) { batch_release(b); return r; }
        char* ring = static_cast<char*>(c->pinned);
        std::atomic<size_t> next{0};
        std::atomic<int> cuda_err{(int)cudaSuccess};
        auto worker = [&](unsigned t) {
            cudaSetDevice(c->device);
            cudaEvent_t ev[2] = {nullptr, nullptr};
            bool used[2] = {false, false};
            cudaError_t err = cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
            if (err == cudaSuccess) err = cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
            int slot = 0;
            while (err == cudaSuccess && cuda_err.load(std::memory_order_relaxed) == (int)cudaSuccess) {
                const size_t i = next.fetch_add(1, std::memory_order_relaxed);
                if (i >= pieces.size()) break;
                const Piece& pc = pieces[i];
                char* stage = ring + ((size_t)2 * t + slot) * PIECE;
                if (used[slot]) err = cudaEventSynchronize(ev[slot]);      // its previous copy has left
                if (err != cudaSuccess) break;
                memcpy(stage, seqs[pc.seq] + pc.off, pc.bytes);
                err = cudaMemcpyAsync(b->chars + b->h_coff[pc.seq] + pc.off, stage, pc.bytes, cudaMemcpyHostToDevice, st);
                if (err == cudaSuccess) err = cudaEventRecord(ev[slot], st);
                used[slot] = true;
                slot ^= 1;
            }
            for (int s2 = 0; s2 < 2; ++s2) {
                if (used[s2] && err == cudaSuccess) err = cudaEventSynchronize(ev[s2]);
                if (ev[s2]) cudaEventDestroy(ev[s2]);
            }
            if (err != cudaSuccess) cuda_err.store((int)err, std::memory_order_relaxed);
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < team; ++t) th.emplace_back(worker, t);
        worker(0);
        for (auto& x : th) x.join();
        e = (cudaError_t)cuda_err.load();
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // staging is reused by the next call
        if (e != cudaSuccess) { batch_release(b); return fail(c, MK_ERR_CUDA, cudaGetErrorString(e)); }
    } else if (n) {
        // many short sequences (reads): pack into pinned staging, one copy
        r = reserve_pinned(c, b->h_coff[n] + 64);
        if (r != MK_OK) { batch_release(b); return r; }
        char* pin = static_cast<char*>(c->pinned);
        for (uint32_t i = 0; i < n; ++i) {
            memcpy(pin + b->h_coff[i], seqs[i], lens[i]);
            const uint64_t pad = b->h_coff[i + 1] - b->h_coff[i] - lens[i];
            if (pad) memset(pin + b->h_coff[i] + lens[i], 0, pad);
        }
        e = cudaMemcpyAsync(b->chars, pin, b->h_coff[n], cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { batch_release(b); return fail(c, MK_ERR_CUDA, cudaGetErrorString(e)); }
        // the staging buffer is reused by the next call: wait for the copy
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { batch_release(b); return fail(c, MK_ERR_CUDA, cudaGetErrorString(e)); }
    }
    c->h2d_meta.fetch_add(b->bases, std::memory_order_relaxed);
    *out = b;
    return MK_OK;
}

// a view on sequences [first, first+n) of an existing batch (shares chars)
struct BatchView {
    const uint8_t* chars;
    const uint64_t* d_coff;
    const uint64_t* d_len;
    const uint64_t* h_len;
    uint32_t n;
    uint64_t max_len, bases;
    // sequences packed on the host (pack.h) instead of chars / d_coff
    const uint32_t* packed = nullptr;   // word w of sequence s at packed[woff(s) + w], woff as in dense_sketch
    const PackExcDev* exc = nullptr;
    uint32_t n_exc = 0, seq0 = 0;
    const uint8_t* pvalid = nullptr;
};
BatchView view_of(const mk_batch* b, uint32_t first, uint32_t n) {
    BatchView v{b->chars, b->d_coff + first, b->d_len + first, b->h_len.data() + first, n, 0, 0};
    for (uint32_t i = 0; i < n; ++i) {
        v.max_len = std::max(v.max_len, v.h_len[i]);
        v.bases += v.h_len[i];
    }
    return v;
}

// ---- dense sketch of a view: planes + keys -> fp / anc ------------------------------
// leaves: c->keys = anc[n][B], c->fp = fp[n][B], c->meta = {woff[n+1] | active[n] | ssum[n]}
struct DenseOut {
    uint64_t* d_woff;
    uint32_t* d_active;
    unsigned long long* d_ssum;
    uint32_t* d_claims;       // Bloom claimants of this chunk (seq << h | bucket), count in *d_n_claims
    uint32_t* d_n_claims;
};
int dense_sketch(mk_ctx* c, const BatchView& v, bool bloom_insert, DenseOut* out) {
    const uint32_t n = v.n;
    std::vector<uint64_t> woff((size_t)n + 1);
    uint64_t w = 0;
    for (uint32_t i = 0; i < n; ++i) {
        woff[i] = w;
        w += (v.h_len[i] + 15) / 16 + 2;
    }
    woff[n] = w;
    TRY(reserve(c, c->planeF, (w + 4) * 8));      // F and R words interleaved: {F, R} per 16 bases
    TRY(reserve(c, c->keys, (size_t)n * c->B * 8));
    const size_t fp_bytes = ((size_t)n * c->B + 3) & ~(size_t)3;
    TRY(reserve(c, c->fp, fp_bytes + (bloom_insert ? (size_t)n * c->B * 4 : 0)));   // fp[n][B] | claim list u32[n * B]
    // woff[n+1] | ssum[n] | active[n] | claim counter  (ssum .. counter zeroed together below)
    const size_t meta_bytes = ((size_t)n + 1) * 8 + (size_t)n * 8 + (size_t)n * 8 + 8;
    TRY(reserve(c, c->meta, meta_bytes));
    uint64_t* d_woff = static_cast<uint64_t*>(c->meta.p);
    unsigned long long* d_ssum = reinterpret_cast<unsigned long long*>(d_woff + n + 1);
    uint32_t* d_active = reinterpret_cast<uint32_t*>(d_ssum + n);
    CU(cudaMemcpyAsync(d_woff, woff.data(), ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    uint32_t* d_n_claims = d_active + n;
    uint32_t* d_claims = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(c->fp.p) + fp_bytes);
    CU(cudaMemsetAsync(d_ssum, 0, (size_t)n * 12 + 4, c->stream));
    auto* keys = static_cast<unsigned long long*>(c->keys.p);
    // tagged keys whenever positions fit 32 bits (common.cuh: KEY_TAG_BITS)
    const int ks = (c->key_tags && v.max_len < (1ull << 32)) ? KEY_TAG_BITS : 0;
    const uint32_t* pair_full = nullptr;
    uint32_t pair_words = 0;
    if (bloom_insert && ks) {
        // which Bloom pages are already saturated (resolve_kernel's fast path)
        const uint32_t n_pages = bloom_page_count(c->window);
        const size_t off = ((size_t)n_pages + 3) & ~(size_t)3;
        TRY(reserve(c, c->pages, off + ((size_t)n_pages + 31) / 32 * 4));
        auto* full8 = static_cast<uint8_t*>(c->pages.p);
        auto* pf = reinterpret_cast<uint32_t*>(full8 + off);
        pair_words = launch_bloom_pages(c->bloom, c->window, c->bloom_reach, full8, pf, c->stream);
        pair_full = pf;
        c->stats.kernel_launches += 2;
    }
    launch_fill_u64(keys, (uint64_t)n * c->B, ~0ull, c->stream);
    if (v.packed)
        launch_expand_planes(v.packed, v.exc, v.n_exc, v.seq0, v.pvalid, v.d_len, d_woff, n, v.max_len, (int)c->k,
                             static_cast<uint32_t*>(c->planeF.p), c->stream);
    else
        launch_encode_planes(v.chars, v.d_coff, v.d_len, d_woff, n, v.max_len, (int)c->k,
                             static_cast<uint32_t*>(c->planeF.p), static_cast<uint32_t*>(c->planeF.p) + 1, c->stream);
    if (v.packed && v.n_exc) c->stats.kernel_launches += 1;
    launch_sketch_dense(static_cast<uint32_t*>(c->planeF.p), static_cast<uint32_t*>(c->planeF.p) + 1, v.d_len,
                        d_woff, n, v.max_len, (int)c->k, (int)c->h, keys, ks, c->stream);
    static const bool split = [] {
        const char* e = getenv("MIEKKI_RESOLVE_SPLIT");
        return !e || atoi(e) != 0;
    }();
    if (split && bloom_insert && ks && pair_full && resolve_tagged_ok((int)c->h, pair_words)) {
        launch_resolve_tagged(keys, static_cast<uint32_t*>(c->planeF.p), d_woff, n, c->sp(),
                              static_cast<uint8_t*>(c->fp.p), d_active, d_ssum, c->bloom, c->owner, pair_full, pair_words,
                              ks, d_claims, d_n_claims, c->stream);
        c->stats.kernel_launches += 1;
    } else {
        launch_resolve(keys, static_cast<uint32_t*>(c->planeF.p), static_cast<uint32_t*>(c->planeF.p) + 1, d_woff, n,
                       c->sp(), static_cast<uint8_t*>(c->fp.p), d_active, d_ssum, c->bloom,
                       bloom_insert ? c->owner : nullptr, pair_full, pair_words, ks, d_claims, d_n_claims, c->stream);
    }
    c->stats.kernel_launches += 4;
    CU(cudaGetLastError());
    out->d_woff = d_woff;
    out->d_active = d_active;
    out->d_ssum = d_ssum;
    out->d_claims = d_claims;
    out->d_n_claims = d_n_claims;
    return MK_OK;
}

// genomes per dense chunk: bounded by the Bloom owner key (6 + h + 3 bits <= 32) and memory.
// 64 genomes are two whole 32-genome groups, i.e. full 32-byte sectors of a row per scatter.
uint32_t dense_chunk(const mk_ctx* c) {
    static const uint32_t want = [] {
        const char* e = getenv("MIEKKI_BUILD_CHUNK");
        return (uint32_t)std::max(1, std::min(64, e ? atoi(e) : 64));
    }();
    uint32_t by_key = c->h <= 24 ? (1u << std::min<uint32_t>(29 - c->h, 6)) : 1;
    uint64_t by_mem = (2ull << 30) / (c->B * 13);       // keys + fp + claim list <= 2 GiB
    uint32_t ch = (uint32_t)std::min<uint64_t>(by_key, std::max<uint64_t>(1, by_mem));
    ch = std::max<uint32_t>(1, std::min<uint32_t>(ch, want));
    return ch >= 32 ? ch / 32 * 32 : ch;                // whole groups keep every chunk group-aligned
}

// Miekki.cpp:303-311 on the host, from the device's exact integer statistics
uint64_t genome_size_of(uint32_t active, unsigned long long ssum31, uint64_t len) {
    const double S = std::ldexp((double)ssum31, -31);            // sum of 2^-(fp>>3), exact
    const uint32_t sq = active * active;                          // uint32 wrap: quirk G2
    const double card = (0.72134 * (double)sq) / S;
    if (card > (double)len) return len;
    return (uint64_t)card;
}

// completes the host mirrors of sketch_size / genome_size up to c->n
int host_stats(mk_ctx* c) {
    const size_t have = c->h_sketch_size.size();
    if (have >= c->n) return MK_OK;
    const size_t m = c->n - have;
    c->h_sketch_size.resize(c->n);
    c->h_genome_size.resize(c->n);
    CU(cudaMemcpyAsync(c->h_sketch_size.data() + have, c->d_sketch_size + have, m * 4, cudaMemcpyDeviceToHost,
                       c->stream));
    CU(cudaMemcpyAsync(c->h_genome_size.data() + have, c->d_genome_size + have, m * 8, cudaMemcpyDeviceToHost,
                       c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->stats.d2h_bytes += m * 12;
    return MK_OK;
}

// Genomes packed on the host (pack.h): 0.25 B per base over PCIe instead of 1 B.
struct PackedBatch {
    uint32_t* d_words = nullptr;     // one allocation: words | len[n] | pvalid[n]
    uint64_t* d_len = nullptr;
    uint8_t* d_pvalid = nullptr;
    PackExcDev* d_exc = nullptr;     // second allocation (its size is known after packing)
    cudaStream_t stream = nullptr;
    std::vector<uint64_t> h_len, h_woff;     // h_woff: n + 1, words per sequence = ceil(len / 16) + 2
    std::vector<uint32_t> h_exc_off;         // n + 1: exceptions of sequence i
    uint32_t n = 0, n_exc = 0;
    uint64_t bases = 0;
};
void packed_release(PackedBatch* pb) {
    if (!pb) return;
    if (pb->d_words) cudaFreeAsync(pb->d_words, pb->stream);
    if (pb->d_exc) cudaFreeAsync(pb->d_exc, pb->stream);
    delete pb;
}
BatchView view_of(const PackedBatch* pb, uint32_t first, uint32_t n) {
    BatchView v{nullptr, nullptr, pb->d_len + first, pb->h_len.data() + first, n, 0, 0};
    for (uint32_t i = 0; i < n; ++i) {
        v.max_len = std::max(v.max_len, v.h_len[i]);
        v.bases += v.h_len[i];
    }
    v.packed = pb->d_words + pb->h_woff[first];
    v.exc = pb->d_exc + pb->h_exc_off[first];
    v.n_exc = pb->h_exc_off[first + n] - pb->h_exc_off[first];
    v.seq0 = first;
    v.pvalid = pb->d_pvalid + first;
    return v;
}

// Packs sequences [0, n) on a team of host threads straight into pinned staging and copies the
// pieces over as they are ready.  Returns MK_OK with *out == nullptr when packing does not pay
// (more than one word in eight needs the general encoder: lower-case or N-rich input): the
// caller then uploads the characters as they are.
int upload_packed(mk_ctx* c, const char* const* seqs, const uint64_t* lens, uint32_t n, PackedBatch** out,
                  cudaStream_t st) {
    *out = nullptr;
    PackedBatch* pb = new PackedBatch();
    pb->n = n;
    pb->stream = st;
    pb->h_len.assign(lens, lens + n);
    pb->h_woff.resize((size_t)n + 1);
    uint64_t words = 0;
    for (uint32_t i = 0; i < n; ++i) {
        pb->h_woff[i] = words;
        words += (lens[i] + 15) / 16 + 2;
        pb->bases += lens[i];
    }
    pb->h_woff[n] = words;
    const size_t words_bytes = (words * 4 + 255) / 256 * 256;
    const size_t total = words_bytes + (size_t)n * 8 + std::max<size_t>(n, 1);
    auto bail = [&](cudaError_t e) {
        packed_release(pb);
        return fail(c, e == cudaErrorMemoryAllocation ? MK_ERR_NOMEM : MK_ERR_CUDA, std::string("packed upload: ") + cudaGetErrorString(e));
    };
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&pb->d_words), total, st);
    if (e != cudaSuccess) return bail(e);
    pb->d_len = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(pb->d_words) + words_bytes);
    pb->d_pvalid = reinterpret_cast<uint8_t*>(pb->d_len + n);
    // pieces of 512 Ki words (8 MB of characters -> 2 MB packed)
    constexpr uint64_t PIECE_WORDS = 512 << 10;
    struct Piece { uint32_t seq; uint64_t w0, w1; };
    std::vector<Piece> pieces;
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t nw = (lens[i] + 15) / 16;
        for (uint64_t w = 0; w < nw; w += PIECE_WORDS) pieces.push_back({i, w, std::min(nw, w + PIECE_WORDS)});
    }
    static const unsigned TEAM = [] {
        const char* ev = getenv("MIEKKI_PACK_THREADS");
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        return (unsigned)std::min(64, std::max(1, ev ? atoi(ev) : (int)std::min(32u, hw)));
    }();
    const unsigned team = (unsigned)std::min<size_t>(TEAM, std::max<size_t>(1, pieces.size()));
    int r = reserve_pinned(c, PIECE_WORDS * 4 * 2 * team);
    if (r != MK_OK) { packed_release(pb); return r; }
    uint32_t* ring = static_cast<uint32_t*>(c->pinned);
    std::atomic<size_t> next{0};
    std::atomic<int> cuda_err{(int)cudaSuccess};
    struct Exc { uint32_t seq; PackException x; };
    std::vector<std::vector<Exc>> found(team);
    auto worker = [&](unsigned t) {
        cudaSetDevice(c->device);
        cudaEvent_t ev[2] = {nullptr, nullptr};
        bool used[2] = {false, false};
        cudaError_t err = cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
        if (err == cudaSuccess) err = cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
        std::vector<PackException> exc;
        int slot = 0;
        while (err == cudaSuccess && cuda_err.load(std::memory_order_relaxed) == (int)cudaSuccess) {
            const size_t i = next.fetch_add(1, std::memory_order_relaxed);
            if (i >= pieces.size()) break;
            const Piece& pc = pieces[i];
            uint32_t* stage = ring + ((size_t)2 * t + slot) * PIECE_WORDS;
            if (used[slot]) err = cudaEventSynchronize(ev[slot]);          // its previous copy has left
            if (err != cudaSuccess) break;
            exc.clear();
            pack_words(seqs[pc.seq], lens[pc.seq], pc.w0, pc.w1, (int)c->k, stage, exc);
            for (const PackException& x : exc) found[t].push_back({pc.seq, x});
            err = cudaMemcpyAsync(pb->d_words + pb->h_woff[pc.seq] + pc.w0, stage, (pc.w1 - pc.w0) * 4,
                                  cudaMemcpyHostToDevice, st);
            if (err == cudaSuccess) err = cudaEventRecord(ev[slot], st);
            used[slot] = true;
            slot ^= 1;
        }
        for (int s2 = 0; s2 < 2; ++s2) {
            if (used[s2] && err == cudaSuccess) err = cudaEventSynchronize(ev[s2]);
            if (ev[s2]) cudaEventDestroy(ev[s2]);
        }
        if (err != cudaSuccess) cuda_err.store((int)err, std::memory_order_relaxed);
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < team; ++t) th.emplace_back(worker, t);
    worker(0);
    for (auto& x : th) x.join();
    e = (cudaError_t)cuda_err.load();
    if (e != cudaSuccess) return bail(e);
    // exceptions in (sequence, word) order, with the offsets of every sequence
    std::vector<Exc> all;
    for (auto& v : found) all.insert(all.end(), v.begin(), v.end());
    if (all.size() > words / 8 + 4 * (size_t)n) {                          // packing does not pay here
        cudaStreamSynchronize(st);
        packed_release(pb);
        return MK_OK;
    }
    std::sort(all.begin(), all.end(), [](const Exc& a, const Exc& b) {
        return a.seq != b.seq ? a.seq < b.seq : a.x.word < b.x.word;
    });
    pb->n_exc = (uint32_t)all.size();
    pb->h_exc_off.assign((size_t)n + 1, 0);
    std::vector<PackExcDev> dev(all.size());
    for (size_t i = 0; i < all.size(); ++i) {
        memcpy(&dev[i].bytes, all[i].x.bytes, 16);
        dev[i].word = all[i].x.word;
        dev[i].seq = all[i].seq;
        dev[i].pad = 0;
        pb->h_exc_off[all[i].seq + 1] += 1;
    }
    for (uint32_t i = 0; i < n; ++i) pb->h_exc_off[i + 1] += pb->h_exc_off[i];
    std::vector<uint8_t> pvalid(std::max<size_t>(n, 1));
    for (uint32_t i = 0; i < n; ++i) pvalid[i] = prefix_is_acgt(seqs[i], lens[i], (int)c->k) ? 1 : 0;
    if (!dev.empty()) {
        e = cudaMallocAsync(reinterpret_cast<void**>(&pb->d_exc), dev.size() * sizeof(PackExcDev), st);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(pb->d_exc, dev.data(), dev.size() * sizeof(PackExcDev), cudaMemcpyHostToDevice, st);
    }
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(pb->d_len, pb->h_len.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(pb->d_pvalid, pvalid.data(), n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);        // host vectors and the staging ring are reused
    if (e != cudaSuccess) return bail(e);
    c->h2d_meta.fetch_add(words * 4 + dev.size() * sizeof(PackExcDev) + (size_t)n * 9, std::memory_order_relaxed);
    *out = pb;
    return MK_OK;
}

// appends the sequences of a batch (characters in HBM) or of a packed batch to the index
template <class Batch>
int index_add_any(mk_ctx* c, const Batch* b) {
    c->prep.serial = 0;        // lists sketched ahead saw the old index
    for (uint32_t i = 0; i < b->n; ++i)
        if (b->h_len[i] < c->k)
            return fail(c, MK_ERR_ARG, "mk_index_add: sequence shorter than k (the reference's callers skip these, Miekki.cpp:569)");
    const uint32_t chunk = dense_chunk(c);
    TRY(ensure_capacity(c, c->n + b->n));
    for (uint32_t first = 0; first < b->n; first += chunk) {
        const uint32_t n = std::min(chunk, b->n - first);
        BatchView v = view_of(b, first, n);
        DenseOut d{};
        {
            PhaseTimer t(c, PH_SKETCH);
            TRY(dense_sketch(c, v, true, &d));
            launch_bloom_commit(static_cast<unsigned long long*>(c->keys.p), d.d_claims, d.d_n_claims, n, c->sp(),
                                c->bloom, c->owner, c->stream);
            launch_scatter_planes(static_cast<uint8_t*>(c->fp.p), n, (int)c->h, c->rows, c->stride, c->n, c->stream);
            c->stats.kernel_launches += 2;
        }
        if (getenv("MIEKKI_TRACE_BUILD")) {      // debugging aid (synchronises): buckets that took the look-up path
            uint32_t slow = 0;
            CU(cudaStreamSynchronize(c->stream));
            CU(cudaMemcpy(&slow, d.d_n_claims, 4, cudaMemcpyDeviceToHost));
            fprintf(stderr, "[trace] chunk of %u genomes at id %u: %u of %llu buckets looked up for the Bloom table\n", n,
                    c->n, slow, (unsigned long long)n * c->B);
        }
        // sketch_size / genome_size / ratio of the new ids, on the device (no host round trip)
        launch_stats_finalize(d.d_active, d.d_ssum, v.d_len, n, c->d_sketch_size + c->n, c->d_genome_size + c->n,
                              c->d_ratio + c->n, c->stream);
        c->stats.kernel_launches += 1;
        CU(cudaGetLastError());
        c->n += n;
        c->stats.bases_sketched += v.bases;
        c->stats.h2d_bytes += ((size_t)n + 1) * 8;
    }
    return sync(c);
}
int index_add_view(mk_ctx* c, const mk_batch* b) { return index_add_any(c, b); }

// ---- query: reads -> (bucket << 8 | fp) lists ----------------------------------------
struct Lists {
    uint32_t* list;
    uint64_t* list_off;     // device, n+1
    uint32_t* list_len;     // device, n
    // tiled scan (scan_tiled.cu): the same lists sorted by bucket, each ending in sentinels
    void* slist = nullptr;
    uint64_t* soff = nullptr;   // device, n
    bool long_lists = false;    // some list may exceed the tiled kernel's 14 counter planes
};

// Which scan kernel serves a batch of reads?  The tiled kernel streams every row of the index once
// per tile of reads instead of one row per (read, bucket) pair: worth it when the reads of a tile
// together cover the bucket space a few times over (long reads or whole genomes at small -h) and
// there are enough (read tile, genome tile) items to fill the GPU; legal when the list sort fits
// shared memory (-h <= 17).  MIEKKI_SCAN_TILED=0 / 1 forces the choice where legal (tests, A/B
// measurements).
bool want_tiled(const mk_ctx* c, const uint64_t* lens, uint32_t n) {
    if (c->n == 0 || c->h > (uint32_t)TILED_MAX_H || c->h < 5 || n == 0 || !tiled_scan_available()) return false;
    const char* e = getenv("MIEKKI_SCAN_TILED");
    if (e && atoi(e) == 0) return false;
    if (e && atoi(e) == 1) return true;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; ++i) total += std::min<uint64_t>(lens[i] > c->k ? lens[i] - c->k : 0, c->B);
    TiledPlan plan{};
    if (tiled_plan(c->n, (int)c->h, c->sm_count, c->smem_optin, &plan) != 0) return false;
    // expected (read, row) pairs per staged row: 0.9 = share of k-mers that end up as list entries
    const double cover = 0.9 * (double)plan.tile_reads * ((double)total / n) / (double)c->B;
    const uint64_t items = (uint64_t)((n + plan.tile_reads - 1) / plan.tile_reads) * plan.n_gt;
    return cover >= 2.0 && c->n >= 768 && items >= 2 * (uint64_t)c->sm_count;
}

constexpr uint64_t SPARSE_MAX_KMERS = 12288;   // 16384-slot table at load <= 0.75

// small pinned staging areas for per-call metadata (two, alternating): uploads from them never
// make the host wait for work already queued on the stream, which mk_scan_async relies on
int pinned_meta(mk_ctx* c, size_t bytes, void** out) {
    const int w = (c->meta_flip ^= 1);
    // the copy that last read this area must have run before the host overwrites it
    if (c->meta_ev[w]) CU(cudaEventSynchronize(c->meta_ev[w]));
    else CU(cudaEventCreateWithFlags(&c->meta_ev[w], cudaEventDisableTiming));
    if (bytes > c->meta_pin_cap[w]) {
        CU(cudaStreamSynchronize(c->stream));
        if (c->meta_pin[w]) CU(cudaFreeHost(c->meta_pin[w]));
        c->meta_pin[w] = nullptr;
        c->meta_pin_cap[w] = 0;
        CU(cudaMallocHost(&c->meta_pin[w], bytes + bytes / 4 + 4096));
        c->meta_pin_cap[w] = bytes + bytes / 4 + 4096;
    }
    *out = c->meta_pin[w];
    return MK_OK;
}

// lists of reads [first, first + n) of a batch; all indices inside are relative to `first`
// c->bl_stream / c->bl_set (set by the caller, default main stream / set 0) choose the stream the
// sketch runs on and which of the two list buffers it fills, so that query_device can sketch the
// next tile of reads beside the running scan.
int build_lists(mk_ctx* c, const mk_batch* b, uint32_t first, uint32_t n, Lists* out) {
    cudaStream_t st = c->bl_stream ? c->bl_stream : c->stream;
    DevBuf& list_buf = c->bl_set ? c->list2 : c->list;
    DevBuf& len_buf = c->bl_set ? c->list_len2 : c->list_len;
    const uint64_t* b_len = b->h_len.data() + first;
    const uint64_t* b_coff = b->h_coff.data() + first;
    const uint64_t* bd_coff = b->d_coff + first;
    const uint64_t* bd_len = b->d_len + first;
    void* pin = nullptr;
    TRY(pinned_meta(c, ((size_t)n + 1) * 8 + (size_t)n * 4 + 64, &pin));
    uint64_t* off = static_cast<uint64_t*>(pin);
    uint32_t* ids = reinterpret_cast<uint32_t*>(off + n + 1);
    std::vector<uint32_t> sparse_ids, dense_ids;
    uint64_t total = 0, sparse_max = 0;
    for (uint32_t i = 0; i < n; ++i) {
        off[i] = total;
        const uint64_t len = b_len[i];
        const uint64_t nk = len > c->k ? len - c->k : 0;
        // a list can hold at most one entry per k-mer and per bucket
        total += std::min<uint64_t>(nk, c->B);
        if (nk == 0) continue;                       // empty sketch (len <= k), empty list
        if (nk <= SPARSE_MAX_KMERS) {
            sparse_ids.push_back(i);
            sparse_max = std::max(sparse_max, len);
        } else {
            dense_ids.push_back(i);
        }
    }
    off[n] = total;
    TRY(reserve(c, list_buf, (total + 4) * 4));
    // list_len | list_off | read ids, all in one small buffer
    const size_t ll_bytes = ((size_t)n * 4 + 15) / 16 * 16;
    const size_t lo_bytes = ((size_t)n + 1) * 8;
    const size_t id_bytes = (size_t)n * 4 + 16;
    TRY(reserve(c, len_buf, ll_bytes + lo_bytes + id_bytes));
    uint8_t* base = static_cast<uint8_t*>(len_buf.p);
    uint32_t* d_len = reinterpret_cast<uint32_t*>(base);
    uint64_t* d_off = reinterpret_cast<uint64_t*>(base + ll_bytes);
    uint32_t* d_ids = reinterpret_cast<uint32_t*>(base + ll_bytes + lo_bytes);
    CU(cudaMemsetAsync(d_len, 0, ll_bytes, st));
    CU(cudaMemcpyAsync(d_off, off, lo_bytes, cudaMemcpyHostToDevice, st));
    c->stats.h2d_bytes += lo_bytes;
    const size_t n_ids = sparse_ids.size() + dense_ids.size();
    std::copy(sparse_ids.begin(), sparse_ids.end(), ids);
    std::copy(dense_ids.begin(), dense_ids.end(), ids + sparse_ids.size());
    if (n_ids) {
        CU(cudaMemcpyAsync(d_ids, ids, n_ids * 4, cudaMemcpyHostToDevice, st));
        c->stats.h2d_bytes += n_ids * 4;
    }
    CU(cudaEventRecord(c->meta_ev[c->meta_flip], st));   // pinned area free again after this
    auto* list = static_cast<uint32_t*>(list_buf.p);
    PhaseTimer t(c, PH_READ_SKETCH, st);
    if (!sparse_ids.empty()) {
        int r = launch_sketch_reads(b->chars, bd_coff, bd_len, d_ids, (uint32_t)sparse_ids.size(),
                                    sparse_max, c->sp(), c->bloom, d_off, list, d_len, st);
        if (r != 0) return fail(c, MK_ERR_CUDA, "sketch_reads launch configuration failed");
        c->stats.kernel_launches += 1;
        CU(cudaGetLastError());
    }
    if (!dense_ids.empty()) {
        // long reads: dense sketch of gathered views, a few at a time (main stream only: the
        // dense scratch is shared with the build path)
        if (st != c->stream) return fail(c, MK_ERR_STATE, "internal: long reads must be sketched on the main stream");
        const uint32_t chunk = dense_chunk(c);
        for (size_t first = 0; first < dense_ids.size(); first += chunk) {
            const uint32_t m = (uint32_t)std::min<size_t>(chunk, dense_ids.size() - first);
            // gather per-read offsets/lengths for this chunk
            std::vector<uint64_t> coff(m), len(m);
            uint64_t max_len = 0, bases = 0;
            for (uint32_t i = 0; i < m; ++i) {
                const uint32_t rid = dense_ids[first + i];
                coff[i] = b_coff[rid];
                len[i] = b_len[rid];
                max_len = std::max(max_len, len[i]);
                bases += len[i];
            }
            TRY(reserve(c, c->misc, (size_t)m * 16));
            uint64_t* d_c = static_cast<uint64_t*>(c->misc.p);
            uint64_t* d_l = d_c + m;
            CU(cudaMemcpyAsync(d_c, coff.data(), (size_t)m * 8, cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemcpyAsync(d_l, len.data(), (size_t)m * 8, cudaMemcpyHostToDevice, c->stream));
            CU(cudaStreamSynchronize(c->stream));   // coff/len are stack vectors
            BatchView v{b->chars, d_c, d_l, len.data(), m, max_len, bases};
            DenseOut d{};
            TRY(dense_sketch(c, v, false, &d));
            launch_compact_list(static_cast<unsigned long long*>(c->keys.p), static_cast<uint8_t*>(c->fp.p), m,
                                c->sp(), c->bloom, d_ids + sparse_ids.size() + first, d_off, list, d_len,
                                c->stream);
            c->stats.kernel_launches += 1;
            CU(cudaGetLastError());
        }
    }
    out->list = list;
    out->list_off = d_off;
    out->list_len = d_len;
    out->slist = nullptr;
    out->soff = nullptr;
    if (want_tiled(c, b_len, n)) {
        // sorted copies of the lists for the tiled scan: [128 sentinels | list 0 | list 1 | ...]
        DevBuf& sbuf = c->bl_set ? c->slist2 : c->slist;
        void* pin2 = nullptr;
        TRY(pinned_meta(c, (size_t)n * 8 + 64, &pin2));
        uint64_t* soff = static_cast<uint64_t*>(pin2);
        uint64_t at = 128;
        for (uint32_t i = 0; i < n; ++i) {
            soff[i] = at;
            at += sorted_list_capacity(std::min<uint64_t>(b_len[i] > c->k ? b_len[i] - c->k : 0, c->B));
        }
        if (at < (1ull << 32)) {                   // the scan addresses entries with 32 bits
            const size_t list_bytes = at * SORTED_ENTRY_BYTES;
            TRY(reserve(c, sbuf, list_bytes + (size_t)n * 8 + 64));
            void* d_slist = sbuf.p;
            uint64_t* d_soff = reinterpret_cast<uint64_t*>(static_cast<uint8_t*>(sbuf.p) + ((list_bytes + 15) & ~(size_t)15));
            CU(cudaMemcpyAsync(d_soff, soff, (size_t)n * 8, cudaMemcpyHostToDevice, st));
            CU(cudaEventRecord(c->meta_ev[c->meta_flip], st));
            c->stats.h2d_bytes += (size_t)n * 8;
            if (launch_sort_lists(list, d_off, d_len, n, (int)c->h, d_slist, d_soff, st) != 0)
                return fail(c, MK_ERR_CUDA, "sort_lists launch configuration failed (shared-memory opt-in)");
            c->stats.kernel_launches += 2;
            CU(cudaGetLastError());
            out->slist = d_slist;
            out->soff = d_soff;
            out->long_lists = false;
            for (uint32_t i = 0; i < n && !out->long_lists; ++i)
                out->long_lists = std::min<uint64_t>(b_len[i] > c->k ? b_len[i] - c->k : 0, c->B) > TILED_SHORT_ENTRIES;
        }
    }
    for (uint32_t i = 0; i < n; ++i) c->stats.bases_queried += b_len[i];
    return MK_OK;
}

// reads per scan launch: bounded by the count tile buffer
uint32_t scan_batch_reads(const mk_ctx* c, uint32_t n_reads) {
    const uint64_t n_pad = (c->n + 31) / 32 * 32;
    uint64_t budget = 1ull << 30;
    if (const char* e = getenv("MIEKKI_COUNT_TILE_BYTES")) budget = std::max<uint64_t>(1, strtoull(e, nullptr, 10));   // tests: many tiles
    uint64_t q = budget / (n_pad * 4);
    q = std::max<uint64_t>(1, std::min<uint64_t>(q, n_reads));
    return (uint32_t)q;
}

// scans reads [q0, q0+nq) into `counts` (default: c->counts)
int scan_reads(mk_ctx* c, const Lists& L, uint32_t q0, uint32_t nq, const ScanPlan& plan,
               uint32_t* counts = nullptr) {
    PhaseTimer t(c, PH_SCAN);
    CU(cudaMemsetAsync(c->d_work, 0, 4, c->stream));
    uint32_t* out = counts ? counts : static_cast<uint32_t*>(c->counts.p);
    int r;
    if (L.slist) {
        TiledPlan tp{};
        r = tiled_plan(c->n, (int)c->h, c->sm_count, c->smem_optin, &tp);
        if (r == 0)
            r = launch_scan_tiled(tp, c->rows, c->stride, c->n, (int)c->h, L.slist, L.soff + q0, nq, L.long_lists, out,
                                  c->d_work, c->stream);
    } else {
        r = launch_scan(plan, c->rows, c->stride, c->n, L.list, L.list_off + q0, L.list_len + q0, nq, out, c->d_work,
                        c->stream);
    }
    if (r != 0) return fail(c, MK_ERR_CUDA, "scan launch configuration failed");
    c->stats.kernel_launches += 1;
    c->stats.scan_launches += 1;
    CU(cudaGetLastError());
    return MK_OK;
}

// accumulates A(q) statistics from the device list lengths
int account_lists(mk_ctx* c, const Lists& L, uint32_t n, uint32_t* surviving, cudaStream_t st = nullptr) {
    // device-side counters (fetched by mk_stats_get): no host round trip on the async path
    launch_account_rows(L.list_len, n, c->n, c->d_stat, st ? st : c->stream);
    c->stats.kernel_launches += 1;
    if (surviving) {
        if (n) CU(cudaMemcpyAsync(surviving, L.list_len, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->stats.d2h_bytes += (size_t)n * 4;
    }
    return MK_OK;
}

// full query of a device batch; heap state in/out through host buffers (may be NULL)
int query_device(mk_ctx* c, const mk_batch* b, uint32_t K, uint32_t min_score, double min_int,
                 mk_hit* heap_io, uint32_t* len_io, bool chain_in, int finalize) {
    if (K < 1 || K > 64) return fail(c, MK_ERR_ARG, "nresults must be in [1, 64]");
    const uint32_t n = b->n;
    if (n == 0) return MK_OK;
    TRY(reserve(c, c->heap, (size_t)n * K * sizeof(HitDev)));
    TRY(reserve(c, c->heap_len, (size_t)n * 4));
    auto* d_heap = static_cast<HitDev*>(c->heap.p);
    auto* d_hlen = static_cast<uint32_t*>(c->heap_len.p);
    if (chain_in) {
        if (!heap_io || !len_io) return fail(c, MK_ERR_ARG, "mk_query_chain needs heap_io and len_io");
        CU(cudaMemcpyAsync(d_heap, heap_io, (size_t)n * K * sizeof(HitDev), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(d_hlen, len_io, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        c->stats.h2d_bytes += (size_t)n * (K * sizeof(HitDev) + 4);
    } else {
        CU(cudaMemsetAsync(d_hlen, 0, (size_t)n * 4, c->stream));
    }
    // Slices of reads bound the list memory (one u32 per k-mer at most): 1M reads of 10 kbp
    // would otherwise ask for 40 GB of lists.  Each slice is sketched once, then scanned in
    // sub-batches into two count tiles: the top-k of tile i runs on the aux stream while the
    // (DRAM-bound, persistent) scan of tile i+1 runs on the main one.
    uint64_t LIST_BUDGET = 768ull << 20;                       // entries (3 GiB)
    if (const char* e = getenv("MIEKKI_LIST_BUDGET_ENTRIES")) LIST_BUDGET = std::max<uint64_t>(1, strtoull(e, nullptr, 10));
    ScanPlan plan{};
    if (c->n > 0 && scan_plan(c->n, c->stride, c->sm_count, c->smem_optin, c->scan_spare_sms, &plan) != 0)
        return fail(c, MK_ERR_CUDA, "no scan plan for this index width");
    const uint64_t n_pad = (c->n + 31) / 32 * 32;
    cudaEvent_t scanned[2] = {get_event(c), get_event(c)}, done[2] = {get_event(c), get_event(c)};
    cudaEvent_t sketched[2] = {get_event(c), get_event(c)}, lists_free[2] = {get_event(c), get_event(c)};
    cudaEvent_t slice_start = get_event(c);
    bool busy[2] = {false, false}, lists_busy[2] = {false, false};
    int rc = MK_OK;
    uint32_t tile_no = 0;
    // event / stream-wait calls of the pipeline: the first failure ends the query with its text
    auto ck = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == MK_OK)
            rc = fail(c, MK_ERR_CUDA, std::string("query pipeline: ") + what + ": " + cudaGetErrorString(e));
        return e == cudaSuccess;
    };
    ck(cudaEventRecord(slice_start, c->stream), "cudaEventRecord");
    ck(cudaEventRecord(lists_free[0], c->stream), "cudaEventRecord");
    for (uint32_t first = 0; first < n && rc == MK_OK;) {
        uint32_t cnt = 0;
        uint64_t entries = 0;
        while (first + cnt < n) {
            const uint64_t len = b->h_len[first + cnt];
            const uint64_t e = std::min<uint64_t>(len > c->k ? len - c->k : 0, c->B);
            if (cnt && entries + e > LIST_BUDGET) break;
            entries += e;
            ++cnt;
        }
        // Tiles of the slice.  Reads that fit the shared-memory sketch are sketched tile by tile on
        // the sketch stream, one tile ahead of the scan (two list buffer sets); a slice with long
        // reads (dense sketch, main-stream scratch) is sketched at once on the main stream.
        bool has_long = false;
        for (uint32_t i = 0; i < cnt && !has_long; ++i)
            has_long = b->h_len[first + i] > c->k + SPARSE_MAX_KMERS;
        const uint32_t qb = c->n > 0 ? std::max<uint32_t>(1, scan_batch_reads(c, cnt) / 2) : cnt;
        if (c->n > 0) {
            rc = reserve(c, c->counts, (size_t)qb * n_pad * 4);
            if (rc == MK_OK) rc = reserve(c, c->counts2, (size_t)qb * n_pad * 4);
            if (rc != MK_OK) break;
        }
        uint32_t* tile[2] = {static_cast<uint32_t*>(c->counts.p), static_cast<uint32_t*>(c->counts2.p)};
        Lists L[2] = {};
        // tile boundaries: a short first tile (its sketch is the only one not hidden by a scan)
        std::vector<uint32_t> cut{0};
        if (!has_long && cnt > qb / 4) cut.push_back(std::max<uint32_t>(1, qb / 8));
        while (cut.back() < cnt) cut.push_back(std::min<uint64_t>(cnt, (uint64_t)cut.back() + qb));
        auto sketch_tile = [&](size_t ti, int set) -> int {          // lists of reads [cut[ti], cut[ti+1])
            const uint32_t q0 = cut[ti];
            const uint32_t nq = cut[ti + 1] - q0;
            if (lists_busy[set]) ck(cudaStreamWaitEvent(c->sk_stream, lists_free[set], 0), "cudaStreamWaitEvent");   // last scan that read it
            c->bl_stream = c->sk_stream;
            c->bl_set = set;
            int r = build_lists(c, b, first + q0, nq, &L[set]);
            if (r == MK_OK) r = account_lists(c, L[set], nq, nullptr, c->sk_stream);
            c->bl_stream = nullptr;
            c->bl_set = 0;
            ck(cudaEventRecord(sketched[set], c->sk_stream), "cudaEventRecord");
            return r != MK_OK ? r : rc;
        };
        if (has_long) {
            ck(cudaStreamWaitEvent(c->stream, lists_free[0], 0), "cudaStreamWaitEvent");
            if (rc == MK_OK) rc = build_lists(c, b, first, cnt, &L[0]);
            if (rc == MK_OK) rc = account_lists(c, L[0], cnt, nullptr);
        } else {
            ck(cudaStreamWaitEvent(c->sk_stream, slice_start, 0), "cudaStreamWaitEvent");   // heap init etc. precede everything
            if (rc == MK_OK) rc = sketch_tile(0, 0);
        }
        if (rc != MK_OK) break;
        for (size_t ti = 0; ti + 1 < cut.size() && rc == MK_OK; ++ti, ++tile_no) {
            const uint32_t q0 = cut[ti];
            const uint32_t nq = cut[ti + 1] - q0;
            const int s = (int)(tile_no & 1);
            const int set = has_long ? 0 : (int)(ti & 1);
            if (!has_long && ti + 2 < cut.size()) {              // sketch the next tile beside this scan
                rc = sketch_tile(ti + 1, set ^ 1);
                if (rc != MK_OK) break;
            }
            if (!has_long) ck(cudaStreamWaitEvent(c->stream, sketched[set], 0), "cudaStreamWaitEvent");
            if (busy[s]) ck(cudaStreamWaitEvent(c->stream, done[s], 0), "cudaStreamWaitEvent");   // count tile s is free again
            if (rc != MK_OK) break;
            if (c->n > 0) {
                Lists view = L[set];
                const uint32_t rel = has_long ? q0 : 0;          // tile lists start at 0, slice lists at q0
                rc = scan_reads(c, view, rel, nq, plan, tile[s]);
                if (rc != MK_OK) break;
            }
            // an empty shard still passes the heap on and, as the last link, sorts it (:396)
            if (c->n > 0 || finalize) {
                ck(cudaEventRecord(scanned[s], c->stream), "cudaEventRecord");
                ck(cudaStreamWaitEvent(c->aux_stream, scanned[s], 0), "cudaStreamWaitEvent");
                {
                    PhaseTimer t(c, PH_TOPK, c->aux_stream);
                    launch_topk(tile[s], nq, c->n, c->first_id, c->d_sketch_size, c->d_genome_size, c->d_ratio, K,
                                min_score, min_int, d_heap + (size_t)(first + q0) * K, d_hlen + first + q0, finalize,
                                c->aux_stream);
                }
                ck(cudaEventRecord(done[s], c->aux_stream), "cudaEventRecord");
                busy[s] = true;
                c->stats.kernel_launches += 1;
            }
            ck(cudaEventRecord(lists_free[set], c->stream), "cudaEventRecord");   // this list set may be refilled
            lists_busy[set] = true;
        }
        first += cnt;
    }
    for (int s = 0; s < 2; ++s)
        if (busy[s]) ck(cudaStreamWaitEvent(c->stream, done[s], 0), "cudaStreamWaitEvent");
    {
        cudaError_t le = cudaGetLastError();
        ck(cudaStreamSynchronize(c->aux_stream), "cudaStreamSynchronize(aux)");
        ck(cudaStreamSynchronize(c->stream), "cudaStreamSynchronize");
        ck(le, "kernel launch");
    }
    cudaStreamSynchronize(c->sk_stream);
    for (int s = 0; s < 2; ++s) {
        c->ev_pool.push_back(scanned[s]);
        c->ev_pool.push_back(done[s]);
        c->ev_pool.push_back(sketched[s]);
        c->ev_pool.push_back(lists_free[s]);
    }
    c->ev_pool.push_back(slice_start);
    if (rc != MK_OK) return rc;
    if (heap_io) {
        CU(cudaMemcpyAsync(heap_io, d_heap, (size_t)n * K * sizeof(HitDev), cudaMemcpyDeviceToHost, c->stream));
        c->stats.d2h_bytes += (size_t)n * K * sizeof(HitDev);
    }
    if (len_io) {
        CU(cudaMemcpyAsync(len_io, d_hlen, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        c->stats.d2h_bytes += (size_t)n * 4;
    }
    return sync(c);
}

// Read sketch (+ list sort) of batch b into the list buffers of `slot`.  Short reads are sketched
// on the sketch stream, so that the work overlaps the scan still running on the main stream, and
// the main stream is made to wait for it (or, with done != nullptr, `done` is recorded instead and
// the caller waits later); long reads need the main-stream dense scratch and are sketched in line.
int sketch_for_slot(mk_ctx* c, const mk_batch* b, int slot, Lists* L, cudaEvent_t done) {
    const uint32_t n = b->n;
    bool has_long = false;
    for (uint32_t i = 0; i < n && !has_long; ++i) has_long = b->h_len[i] > c->k + SPARSE_MAX_KMERS;
    if (has_long) {
        if (done) return 1;                          // not ahead of the scan: the caller falls back
        c->bl_set = slot;
        int r = build_lists(c, b, 0, n, L);
        c->bl_set = 0;
        TRY(r);
        TRY(account_lists(c, *L, n, nullptr));
        return MK_OK;
    }
    if (!c->sk_ev) CU(cudaEventCreateWithFlags(&c->sk_ev, cudaEventDisableTiming));
    if (c->slot_used[slot]) CU(cudaStreamWaitEvent(c->sk_stream, c->slot_ev[slot], 0));   // last scan of this slot
    // No wait on the main stream: it may hold the previous batch's scan, beside which this
    // sketch is meant to run.  The reads themselves are complete: every call that creates a
    // batch returns only after its copies have finished.
    c->bl_stream = c->sk_stream;
    c->bl_set = slot;
    int r = build_lists(c, b, 0, n, L);
    if (r == MK_OK) r = account_lists(c, *L, n, nullptr, c->sk_stream);
    c->bl_stream = nullptr;
    c->bl_set = 0;
    TRY(r);
    if (done) {
        CU(cudaEventRecord(done, c->sk_stream));
    } else {
        CU(cudaEventRecord(c->sk_ev, c->sk_stream));
        CU(cudaStreamWaitEvent(c->stream, c->sk_ev, 0));
    }
    return MK_OK;
}

// mk_sketch_async: the lists of the batch the NEXT mk_scan_async will be given
int sketch_ahead(mk_ctx* c, const mk_batch* b) {
    c->prep.serial = 0;
    if (b->n == 0) return MK_OK;
    const int slot = c->next_slot ^ 1;
    if (!c->slot_ev[0]) {
        CU(cudaEventCreateWithFlags(&c->slot_ev[0], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->slot_ev[1], cudaEventDisableTiming));
    }
    if (!c->prep.ev) CU(cudaEventCreateWithFlags(&c->prep.ev, cudaEventDisableTiming));
    Lists L{};
    const int r = sketch_for_slot(c, b, slot, &L, c->prep.ev);
    if (r == 1) return MK_OK;                   // long reads: mk_scan_async sketches them itself
    TRY(r);
    c->prep.serial = b->serial;
    c->prep.slot = slot;
    c->prep.list = L.list;
    c->prep.list_off = L.list_off;
    c->prep.list_len = L.list_len;
    c->prep.slist = L.slist;
    c->prep.soff = L.soff;
    c->prep.long_lists = L.long_lists;
    return MK_OK;
}

// sketch + scan of a whole batch with the counts kept in HBM (for the chained top-k).
// Asynchronous: everything is enqueued on the ctx stream, the counts go to one of two tiles
// ("slots") so that the top-k of one batch can run (aux stream) while the next batch scans.
int scan_all_async(mk_ctx* c, const mk_batch* b, int* slot_out) {
    const uint32_t n = b->n;
    const int slot = (c->next_slot ^= 1);
    c->slot_reads[slot] = 0;
    if (slot_out) *slot_out = slot;
    c->last_slot = slot;
    if (n == 0) return MK_OK;
    const uint64_t n_pad = (c->n + 31) / 32 * 32;
    if ((uint64_t)n * std::max<uint64_t>(n_pad, 16) * 4 > (16ull << 30))
        return fail(c, MK_ERR_ARG, "mk_scan: count matrix would exceed 16 GiB, split the reads");
    if (!c->slot_ev[0]) {
        CU(cudaEventCreateWithFlags(&c->slot_ev[0], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->slot_ev[1], cudaEventDisableTiming));
    }
    Lists L{};
    if (c->prep.serial == b->serial && c->prep.slot == slot) {
        // mk_sketch_async has built this batch's lists already (on the sketch stream)
        L.list = c->prep.list;
        L.list_off = c->prep.list_off;
        L.list_len = c->prep.list_len;
        L.slist = c->prep.slist;
        L.soff = c->prep.soff;
        L.long_lists = c->prep.long_lists;
        CU(cudaStreamWaitEvent(c->stream, c->prep.ev, 0));
    } else {
        TRY(sketch_for_slot(c, b, slot, &L, nullptr));
    }
    c->prep.serial = 0;
    DevBuf& tile = slot ? c->counts2 : c->counts;
    if (c->n > 0) {
        ScanPlan plan{};
        if (scan_plan(c->n, c->stride, c->sm_count, c->smem_optin, c->scan_spare_sms, &plan) != 0)
            return fail(c, MK_ERR_CUDA, "no scan plan for this index width");
        TRY(reserve(c, tile, (size_t)n * n_pad * 4));
        TRY(scan_reads(c, L, 0, n, plan, static_cast<uint32_t*>(tile.p)));
    }
    CU(cudaEventRecord(c->slot_ev[slot], c->stream));
    c->slot_used[slot] = true;
    c->slot_reads[slot] = n;
    return MK_OK;
}

// bounded-heap step on the counts of `slot`, on the aux stream; waits for that stream only.
// [first, first + count) is the range of the batch's reads to step (heap_io / len_io address
// read `first`): a chain over shards passes the batch on in tiles, so that the shards work on
// different tiles at the same time.
int topk_slot(mk_ctx* c, int slot, uint32_t first, uint32_t count, uint32_t K, uint32_t min_score, double min_int,
              mk_hit* heap_io, uint32_t* len_io, bool chain_in, int finalize) {
    if (K < 1 || K > 64) return fail(c, MK_ERR_ARG, "nresults must be in [1, 64]");
    if (slot < 0 || slot > 1) return fail(c, MK_ERR_ARG, "mk_topk: bad slot");
    const uint32_t total = c->slot_reads[slot];
    if (first >= total) return MK_OK;            // the range is clipped to the batch
    const uint32_t n = std::min(count, total - first);
    if (n == 0) return MK_OK;
    if (!heap_io || !len_io) return fail(c, MK_ERR_ARG, "mk_topk needs heap_io and len_io");
    cudaStream_t st = c->aux_stream;
    // the heap scratch is per slot as well: two batches may be in flight
    DevBuf& hb = slot ? c->heap2 : c->heap;
    DevBuf& lb = slot ? c->heap_len2 : c->heap_len;
    if ((size_t)total * K * sizeof(HitDev) > hb.cap || (size_t)total * 4 > lb.cap) {
        CU(cudaStreamSynchronize(st));
        TRY(reserve(c, hb, (size_t)total * K * sizeof(HitDev)));
        TRY(reserve(c, lb, (size_t)total * 4));
    }
    auto* d_heap = static_cast<HitDev*>(hb.p) + (size_t)first * K;
    auto* d_hlen = static_cast<uint32_t*>(lb.p) + first;
    CU(cudaStreamWaitEvent(st, c->slot_ev[slot], 0));
    if (chain_in) {
        CU(cudaMemcpyAsync(d_heap, heap_io, (size_t)n * K * sizeof(HitDev), cudaMemcpyDefault, st));
        CU(cudaMemcpyAsync(d_hlen, len_io, (size_t)n * 4, cudaMemcpyDefault, st));
    } else {
        CU(cudaMemsetAsync(d_hlen, 0, (size_t)n * 4, st));
    }
    if (c->n > 0 || finalize) {      // an empty shard still passes the heap on / sorts it
        PhaseTimer t(c, PH_TOPK, st);
        const uint64_t n_pad = (c->n + 31) / 32 * 32;
        launch_topk(static_cast<uint32_t*>((slot ? c->counts2 : c->counts).p) + (size_t)first * n_pad, n, c->n,
                    c->first_id, c->d_sketch_size, c->d_genome_size, c->d_ratio, K, min_score, min_int, d_heap,
                    d_hlen, finalize, st);
        c->stats.kernel_launches += 1;
        CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(heap_io, d_heap, (size_t)n * K * sizeof(HitDev), cudaMemcpyDefault, st));
    CU(cudaMemcpyAsync(len_io, d_hlen, (size_t)n * 4, cudaMemcpyDefault, st));
    CU(cudaStreamSynchronize(st));
    return MK_OK;
}

struct Guard {
    std::lock_guard<std::mutex> lk;
    explicit Guard(mk_ctx* c) : lk(c->mu) { cudaSetDevice(c->device); }
};

}  // namespace

// ======================================================================================
extern "C" {

int mk_abi_version(void) { return MK_ABI_VERSION; }

const char* mk_last_error(const mk_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int mk_create(uint32_t k, uint32_t h, uint32_t bits_per_min, uint32_t bits_mantis, uint32_t bloom_log2,
              uint32_t threshold, int device, mk_ctx** out) {
    mk_ctx* c = nullptr;
    if (!out) return fail(c, MK_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (bits_per_min != 8 || bits_mantis != 5)
        return fail(c, MK_ERR_UNSUPPORTED, "not implemented: only 8-bit fingerprints (-f 3) with a 5-bit exponent (Miekki.cpp:236, quirk G12)");
    if (k < 2 || k > 31) return fail(c, MK_ERR_ARG, "k must be in [2, 31] (Miekki.h:76-77)");
    if (h < 1 || h > 24) return fail(c, MK_ERR_ARG, "h must be in [1, 24]");
    if (bloom_log2 < 32 || bloom_log2 > 40)
        return fail(c, MK_ERR_ARG, "bloom_log2 must be in [32, 40]: below 32 the reference indexes out of bounds (Miekki.cpp:124)");
    // MIEKKI_TIMING=1: where start-up time goes (driver initialisation dominates a short run)
    const bool timing = getenv("MIEKKI_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[timing] mk_create: %s %.3f s\n", what, std::chrono::duration<double>(now - t_last).count());
        t_last = now;
    };
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(c, MK_ERR_CUDA, "no CUDA device: miekki_b200 has no CPU fallback");
    }
    lap("driver initialisation (cudaGetDeviceCount)");
    if (device < 0 || device >= ndev) return fail(c, MK_ERR_ARG, "bad device ordinal");
    cudaDeviceProp prop{};
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess)
        return fail(c, MK_ERR_CUDA, "cudaSetDevice failed");
    if (prop.major < 10)
        return fail(c, MK_ERR_UNSUPPORTED, "miekki_b200 is built for sm_100a (B200) only");
    cudaFree(nullptr);
    lap("context (cudaSetDevice, properties)");
    mk_ctx* ctx = new mk_ctx();
    ctx->k = k; ctx->h = h; ctx->nbm = bits_per_min; ctx->nbmant = bits_mantis; ctx->b = bloom_log2;
    ctx->threshold = threshold;
    if (const char* e = getenv("MIEKKI_KEY_TAGS")) ctx->key_tags = atoi(e) != 0;
    ctx->B = 1ull << h;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    c = ctx;
    // Bloom window: canonical k-mer < 4^k, the probe term adds < 1024 (utils.cpp:197-199)
    const uint64_t top = (1ull << (2 * k)) + 1023;
    uint64_t window = ((top >> bloom_log2) >> 3) + 1;
    window = std::min<uint64_t>(window, (1ull << bloom_log2) / 8);
    ctx->window = (window + 15) / 16 * 16;
    ctx->bloom_reach = std::min<uint64_t>(ctx->window, mk_bloom_reach(k, bloom_log2));
    {   // keep freed per-call buffers in the stream-ordered pool instead of returning them
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    // The scan's stream outranks the helper streams: when a scan and the top-k of the previous
    // tile become runnable together, the persistent scan CTAs are placed first and the top-k
    // fills what is left of each SM instead of delaying the scan by its whole duration.
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    cudaError_t e = cudaStreamCreateWithPriority(&ctx->own_stream, cudaStreamNonBlocking, prio_greatest);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->sk_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->bloom, ctx->window);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->owner, ctx->window * 4);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_work, 16);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_stat, 16);
    if (e == cudaSuccess) e = cudaMemset(ctx->d_stat, 0, 16);
    if (e != cudaSuccess) {
        std::string msg = std::string("mk_create: ") + cudaGetErrorString(e);
        mk_destroy(ctx);
        return fail(nullptr, MK_ERR_CUDA, msg);
    }
    ctx->stream = ctx->own_stream;
    cudaMemsetAsync(ctx->bloom, 0, ctx->window, ctx->stream);
    launch_fill_u32(ctx->owner, ctx->window, 0xFFFFFFFFu, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    lap("streams, Bloom table, module load");
    *out = ctx;
    return MK_OK;
}

void mk_destroy(mk_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->aux_stream) cudaStreamSynchronize(c->aux_stream);
    for (DevBuf* b : {&c->planeF, &c->planeR, &c->keys, &c->fp, &c->meta, &c->list, &c->list_len, &c->counts,
                      &c->list2, &c->list_len2, &c->counts2, &c->heap, &c->heap_len, &c->heap2, &c->heap_len2, &c->misc, &c->pages,
                      &c->slist, &c->slist2})
        if (b->p) cudaFree(b->p);
    for (void* p : {(void*)c->rows, (void*)c->d_sketch_size, (void*)c->d_genome_size, (void*)c->d_ratio, (void*)c->bloom,
                    (void*)c->owner, (void*)c->d_work, (void*)c->d_stat})
        if (p) cudaFree(p);
    if (c->pinned) cudaFreeHost(c->pinned);
    for (void* p : c->meta_pin)
        if (p) cudaFreeHost(p);
    for (cudaEvent_t e : c->slot_ev)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->meta_ev)
        if (e) cudaEventDestroy(e);
    if (c->sk_ev) cudaEventDestroy(c->sk_ev);
    for (auto& pe : c->ev_pending) { cudaEventDestroy(pe.a); cudaEventDestroy(pe.b); }
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->sk_stream) cudaStreamDestroy(c->sk_stream);
    delete c;
}

int mk_set_stream(mk_ctx* c, void* cuda_stream) {
    if (!c) return MK_ERR_ARG;
    Guard g(c);
    CU(cudaStreamSynchronize(c->stream));
    resolve_events(c);
    c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->own_stream;
    return MK_OK;
}

int mk_set_scan_spare_sms(mk_ctx* c, int n) {
    if (!c) return MK_ERR_ARG;
    Guard g(c);
    if (n < 0 || n >= c->sm_count) return fail(c, MK_ERR_ARG, "mk_set_scan_spare_sms: out of range");
    c->scan_spare_sms = n;
    return MK_OK;
}

int mk_set_shard(mk_ctx* c, uint32_t first_id) {
    if (!c) return MK_ERR_ARG;
    Guard g(c);
    c->first_id = first_id;
    return MK_OK;
}

int mk_get_params(const mk_ctx* c, uint32_t* k, uint32_t* h, uint32_t* bits_per_min, uint32_t* bits_mantis,
                  uint32_t* bloom_log2, uint32_t* threshold) {
    if (!c) return MK_ERR_ARG;
    if (k) *k = c->k;
    if (h) *h = c->h;
    if (bits_per_min) *bits_per_min = c->nbm;
    if (bits_mantis) *bits_mantis = c->nbmant;
    if (bloom_log2) *bloom_log2 = c->b;
    if (threshold) *threshold = c->threshold;
    return MK_OK;
}

// ---- batches -------------------------------------------------------------------------

int mk_batch_upload(mk_ctx* c, const char* const* seqs, const uint64_t* lens, uint32_t n, mk_batch** out) {
    if (!c || !out || (n && (!seqs || !lens))) return fail(c, MK_ERR_ARG, "mk_batch_upload: NULL argument");
    Guard g(c);
    TRY(upload_range(c, seqs, lens, n, out));
    return sync(c);
}

int mk_batch_upload_flat(mk_ctx* c, const char* data, const uint64_t* offsets, const uint64_t* lens, uint32_t n,
                         mk_batch** out) {
    if (!c || !out || (n && (!data || !offsets || !lens)))
        return fail(c, MK_ERR_ARG, "mk_batch_upload_flat: NULL argument");
    Guard g(c);
    mk_batch* b = new mk_batch();
    batch_layout(b, lens, n);
    // on the copy stream: the upload must not queue behind a scan running on the main stream
    cudaStream_t st = c->copy_stream;
    int r = batch_alloc(c, b, st);
    if (r != MK_OK) { batch_release(b); return r; }
    cudaError_t e = cudaSuccess;
    uint64_t span = 0;
    bool same = n > 0;
    for (uint32_t i = 0; i < n && same; ++i) {
        same = offsets[i] == b->h_coff[i];
        span = offsets[i] + lens[i];
    }
    if (same) {
        // the caller's layout already is ours (every start 16-byte aligned, packed): one copy.
        // Gap bytes come from the caller's buffer; kernels never interpret bytes past a length.
        e = cudaMemsetAsync(b->chars + span, 0, b->bytes - span, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(b->chars, data, span, cudaMemcpyHostToDevice, st);
    } else {
        e = cudaMemsetAsync(b->chars, 0, b->bytes, st);
        for (uint32_t i = 0; i < n && e == cudaSuccess; ++i)
            if (lens[i])
                e = cudaMemcpyAsync(b->chars + b->h_coff[i], data + offsets[i], lens[i], cudaMemcpyHostToDevice,
                                    st);
    }
    if (e != cudaSuccess) { batch_release(b); return fail(c, MK_ERR_CUDA, cudaGetErrorString(e)); }
    c->stats.h2d_bytes += b->bases;
    *out = b;
    CU(cudaStreamSynchronize(st));
    return MK_OK;
}

int mk_batch_synth(mk_ctx* c, uint64_t seed, uint32_t first_g, uint32_t n, uint64_t len, mk_batch** out) {
    if (!c || !out) return fail(c, MK_ERR_ARG, "mk_batch_synth: NULL argument");
    Guard g(c);
    std::vector<uint64_t> lens(n, len);
    mk_batch* b = new mk_batch();
    batch_layout(b, lens.data(), n);
    int r = batch_alloc(c, b);
    if (r != MK_OK) { batch_release(b); return r; }
    cudaMemsetAsync(b->chars + b->h_coff[n], 0, b->bytes - b->h_coff[n], c->stream);
    launch_synth(b->chars, b->d_coff, seed, first_g, n, len, c->stream);
    c->stats.kernel_launches += 1;
    *out = b;
    return sync(c);
}

int mk_batch_download(mk_ctx* c, const mk_batch* b, uint32_t i, char* dst, uint64_t cap) {
    if (!c || !b || !dst || i >= b->n) return fail(c, MK_ERR_ARG, "mk_batch_download: bad argument");
    Guard g(c);
    const uint64_t m = std::min<uint64_t>(cap, b->h_len[i]);
    CU(cudaMemcpyAsync(dst, b->chars + b->h_coff[i], m, cudaMemcpyDeviceToHost, c->stream));
    return sync(c);
}

uint32_t mk_batch_size(const mk_batch* b) { return b ? b->n : 0; }
uint64_t mk_batch_bases(const mk_batch* b) { return b ? b->bases : 0; }

void mk_batch_free(mk_ctx* c, mk_batch* b) {
    if (!b) return;
    if (c) {
        Guard g(c);
        // The memory goes back to the pool in stream order, after everything enqueued so far that
        // may read it (scan lists on the main stream, the read sketch on its own stream): the
        // host does not wait for a running scan.
        bool ordered = b->stream != nullptr;
        for (cudaStream_t user : {c->stream, c->sk_stream}) {
            if (!ordered || user == b->stream || !user) continue;
            cudaEvent_t ev = nullptr;
            ordered = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess &&
                      cudaEventRecord(ev, user) == cudaSuccess &&
                      cudaStreamWaitEvent(b->stream, ev, 0) == cudaSuccess;
            if (ev) cudaEventDestroy(ev);       // released once the recorded work has completed
        }
        if (!ordered) {
            cudaGetLastError();
            cudaStreamSynchronize(c->stream);
            cudaStreamSynchronize(c->sk_stream);
        }
        batch_release(b);
    } else {
        batch_release(b);
    }
}

// ---- build ---------------------------------------------------------------------------

int mk_index_reserve(mk_ctx* c, uint32_t n_genomes) {
    if (!c) return MK_ERR_ARG;
    Guard g(c);
    return ensure_capacity(c, n_genomes);
}

int mk_index_add_batch(mk_ctx* c, const mk_batch* genomes) {
    if (!c || !genomes) return fail(c, MK_ERR_ARG, "mk_index_add_batch: NULL argument");
    Guard g(c);
    return index_add_view(c, genomes);
}

int mk_index_add(mk_ctx* c, const char* const* seqs, const uint64_t* lens, uint32_t n) {
    if (!c || (n && (!seqs || !lens))) return fail(c, MK_ERR_ARG, "mk_index_add: NULL argument");
    Guard g(c);
    for (uint32_t i = 0; i < n; ++i)
        if (lens[i] < c->k) return fail(c, MK_ERR_ARG, "mk_index_add: sequence shorter than k");
    // Slices of one dense chunk (<= 32 genomes, <= 1 GiB): while slice i is sketched on the main
    // stream, a helper thread stages slice i+1 through the pinned ring on the copy stream.
    const uint32_t chunk = dense_chunk(c);
    std::vector<std::pair<uint32_t, uint32_t>> slices;
    for (uint32_t first = 0; first < n;) {
        uint64_t bytes = 0;
        uint32_t m = 0;
        while (first + m < n && m < chunk && (m == 0 || bytes + lens[first + m] <= (1ull << 30))) {
            bytes += lens[first + m];
            ++m;
        }
        slices.push_back({first, m});
        first += m;
    }
    TRY(ensure_capacity(c, c->n + n));
    // Long sequences are packed to 2 bits per base on the host (pack.h) so that a quarter of the
    // bytes cross PCIe; short ones, and input that packs badly, go over as characters.
    // (read per call so that tests can switch: MIEKKI_PACKED_UPLOAD=0 never packs,
    // MIEKKI_PACK_MIN_LEN = smallest mean sequence length that is packed, default 65,536)
    const char* pe = getenv("MIEKKI_PACKED_UPLOAD");
    const bool pack_on = !pe || atoi(pe) != 0;
    const char* pm = getenv("MIEKKI_PACK_MIN_LEN");
    const uint64_t pack_min = pm ? strtoull(pm, nullptr, 10) : (1u << 16);
    struct Staged {
        int r;
        mk_batch* b;
        PackedBatch* pb;
    };
    auto stage = [c, seqs, lens, pack_on, pack_min](uint32_t first, uint32_t m) {
        cudaSetDevice(c->device);
        Staged s{MK_OK, nullptr, nullptr};
        uint64_t bases = 0;
        for (uint32_t i = 0; i < m; ++i) bases += lens[first + i];
        if (pack_on && m && bases / m >= pack_min) s.r = upload_packed(c, seqs + first, lens + first, m, &s.pb, c->copy_stream);
        if (s.r == MK_OK && !s.pb) s.r = upload_range(c, seqs + first, lens + first, m, &s.b, c->copy_stream);
        return s;
    };
    std::future<Staged> next;
    if (!slices.empty()) next = std::async(std::launch::async, stage, slices[0].first, slices[0].second);
    int rc = MK_OK;
    for (size_t i = 0; i < slices.size(); ++i) {
        Staged s = next.get();
        if (i + 1 < slices.size())
            next = std::async(std::launch::async, stage, slices[i + 1].first, slices[i + 1].second);
        int r = s.r;
        if (r == MK_OK && rc == MK_OK) r = s.pb ? index_add_any(c, s.pb) : index_add_view(c, s.b);
        if (s.b || s.pb) {
            cudaStreamSynchronize(c->stream);
            batch_release(s.b);
            packed_release(s.pb);
        }
        if (r != MK_OK && rc == MK_OK) rc = r;
    }
    return rc;
}

int mk_index_size(const mk_ctx* c, uint32_t* n_genomes) {
    if (!c || !n_genomes) return MK_ERR_ARG;
    *n_genomes = c->n;
    return MK_OK;
}

int mk_index_stats(mk_ctx* c, uint32_t first, uint32_t n, uint32_t* sketch_size, uint64_t* genome_size) {
    if (!c) return MK_ERR_ARG;
    Guard g(c);
    if ((uint64_t)first + n > c->n) return fail(c, MK_ERR_ARG, "mk_index_stats: range exceeds the index");
    TRY(host_stats(c));
    if (sketch_size) memcpy(sketch_size, c->h_sketch_size.data() + first, (size_t)n * 4);
    if (genome_size) memcpy(genome_size, c->h_genome_size.data() + first, (size_t)n * 8);
    return MK_OK;
}

// rows [row0, row0 + nrows) of the dump payload: bit-plane rows -> dense byte rows on the
// device, a slab of buckets at a time, then to the caller with its own row pitch
static int export_rows(mk_ctx* c, uint64_t row0, uint64_t nrows, uint8_t* dst, uint64_t dst_stride) {
    if (!nrows || !c->n) return MK_OK;
    const uint64_t slab = std::max<uint64_t>(1, (256ull << 20) / c->n);
    TRY(reserve(c, c->misc, std::min<uint64_t>(slab, nrows) * c->n));
    for (uint64_t r0 = 0; r0 < nrows; r0 += slab) {
        const uint64_t nr = std::min<uint64_t>(slab, nrows - r0);
        launch_planes_to_bytes(c->rows, c->stride, row0 + r0, nr, c->n, static_cast<uint8_t*>(c->misc.p), c->stream);
        CU(cudaMemcpy2DAsync(dst + r0 * dst_stride, dst_stride, c->misc.p, c->n, c->n, nr, cudaMemcpyDeviceToHost,
                             c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->stats.kernel_launches += 1;
    }
    c->stats.d2h_bytes += nrows * c->n;
    return MK_OK;
}

int mk_index_export_rows(mk_ctx* c, uint64_t row0, uint64_t nrows, uint8_t* dst, uint64_t dst_stride) {
    if (!c || (nrows && !dst)) return fail(c, MK_ERR_ARG, "mk_index_export_rows: NULL argument");
    Guard g(c);
    if (row0 + nrows > c->B) return fail(c, MK_ERR_ARG, "mk_index_export_rows: range exceeds 2^h rows");
    if (nrows && dst_stride < c->n) return fail(c, MK_ERR_ARG, "mk_index_export_rows: dst_stride < n");
    return export_rows(c, row0, nrows, dst, dst_stride);
}

int mk_index_export(mk_ctx* c, uint8_t* rows, uint64_t* genome_size, uint8_t* bloom, uint64_t bloom_bytes,
                    uint32_t* sketch_size) {
    if (!c) return MK_ERR_ARG;
    Guard g(c);
    if (rows) TRY(export_rows(c, 0, c->B, rows, c->n));
    if (bloom) {
        const uint64_t m = std::min<uint64_t>(bloom_bytes, c->window);
        CU(cudaMemcpyAsync(bloom, c->bloom, m, cudaMemcpyDeviceToHost, c->stream));
        if (bloom_bytes > m) memset(bloom + m, 0, bloom_bytes - m);   // never touched for this k
        c->stats.d2h_bytes += m;
    }
    if (genome_size || sketch_size) TRY(host_stats(c));
    if (genome_size) memcpy(genome_size, c->h_genome_size.data(), (size_t)c->n * 8);
    if (sketch_size) memcpy(sketch_size, c->h_sketch_size.data(), (size_t)c->n * 4);
    return sync(c);
}

// ---- load: begin(n) -> rows in any number of slabs -> end(statistics, Bloom) -----------------
static int import_begin(mk_ctx* c, uint32_t n) {
    c->n = 0;
    c->importing = 0;
    c->h_sketch_size.clear();
    c->h_genome_size.clear();
    TRY(ensure_capacity(c, std::max<uint32_t>(n, 1)));
    CU(cudaMemsetAsync(c->rows, 0xFF, c->B * c->stride, c->stream));
    c->importing = n;
    c->import_open = true;
    return MK_OK;
}

static int import_rows(mk_ctx* c, uint64_t row0, uint64_t nrows, const uint8_t* src, uint64_t src_stride) {
    const uint32_t n = c->importing;
    if (!n || !nrows) return MK_OK;
    // dense byte rows (dump layout) -> bit-plane rows, a slab of buckets at a time
    const uint64_t slab = std::max<uint64_t>(1, (256ull << 20) / n);
    TRY(reserve(c, c->misc, std::min<uint64_t>(slab, nrows) * n));
    for (uint64_t r0 = 0; r0 < nrows; r0 += slab) {
        const uint64_t nr = std::min<uint64_t>(slab, nrows - r0);
        CU(cudaMemcpy2DAsync(c->misc.p, n, src + r0 * src_stride, src_stride, n, nr, cudaMemcpyHostToDevice, c->stream));
        launch_bytes_to_planes(static_cast<uint8_t*>(c->misc.p), n, row0 + r0, nr, n, c->rows, c->stride, c->stream);
        CU(cudaStreamSynchronize(c->stream));
        c->stats.kernel_launches += 1;
    }
    c->stats.h2d_bytes += nrows * n;
    return MK_OK;
}

static int import_end(mk_ctx* c, const uint64_t* genome_size, const uint8_t* bloom, uint64_t bloom_bytes,
                      const uint32_t* sketch_size) {
    const uint32_t n = c->importing;
    if (n) {
        CU(cudaMemcpyAsync(c->d_sketch_size, sketch_size, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->d_genome_size, genome_size, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
        c->stats.h2d_bytes += (size_t)n * 12;
    }
    CU(cudaMemsetAsync(c->bloom, 0, c->window, c->stream));
    if (bloom && bloom_bytes) {
        const uint64_t m = std::min<uint64_t>(bloom_bytes, c->window);
        CU(cudaMemcpyAsync(c->bloom, bloom, m, cudaMemcpyHostToDevice, c->stream));
        c->stats.h2d_bytes += m;
    }
    CU(cudaStreamSynchronize(c->stream));
    c->h_sketch_size.assign(sketch_size, sketch_size + n);
    c->h_genome_size.assign(genome_size, genome_size + n);
    TRY(upload_ratio(c, 0, n));
    c->n = n;
    c->importing = 0;
    c->import_open = false;
    return MK_OK;
}

int mk_index_import_begin(mk_ctx* c, uint32_t n) {
    if (!c) return MK_ERR_ARG;
    Guard g(c);
    c->prep.serial = 0;        // lists sketched ahead saw the old index
    return import_begin(c, n);
}

int mk_index_import_rows(mk_ctx* c, uint64_t row0, uint64_t nrows, const uint8_t* src, uint64_t src_stride) {
    if (!c || (nrows && !src)) return fail(c, MK_ERR_ARG, "mk_index_import_rows: NULL argument");
    Guard g(c);
    if (!c->import_open) return fail(c, MK_ERR_STATE, "mk_index_import_rows: call mk_index_import_begin first");
    if (row0 + nrows > c->B) return fail(c, MK_ERR_ARG, "mk_index_import_rows: range exceeds 2^h rows");
    if (nrows && src_stride < c->importing) return fail(c, MK_ERR_ARG, "mk_index_import_rows: src_stride < n");
    return import_rows(c, row0, nrows, src, src_stride);
}

int mk_index_import_end(mk_ctx* c, const uint64_t* genome_size, const uint8_t* bloom, uint64_t bloom_bytes,
                        const uint32_t* sketch_size) {
    if (!c) return MK_ERR_ARG;
    Guard g(c);
    c->prep.serial = 0;        // lists sketched ahead saw the old index
    if (!c->import_open) return fail(c, MK_ERR_STATE, "mk_index_import_end: call mk_index_import_begin first");
    if (c->importing && (!genome_size || !sketch_size))
        return fail(c, MK_ERR_ARG, "mk_index_import_end: NULL argument");
    return import_end(c, genome_size, bloom, bloom_bytes, sketch_size);
}

int mk_index_import(mk_ctx* c, uint32_t n, const uint8_t* rows, uint64_t rows_stride, const uint64_t* genome_size,
                    const uint8_t* bloom, uint64_t bloom_bytes, const uint32_t* sketch_size) {
    if (!c || (n && (!rows || !genome_size || !sketch_size)))
        return fail(c, MK_ERR_ARG, "mk_index_import: NULL argument");
    Guard g(c);
    if (n && rows_stride < n) return fail(c, MK_ERR_ARG, "mk_index_import: rows_stride < n");
    TRY(import_begin(c, n));
    TRY(import_rows(c, 0, c->B, rows, rows_stride));
    return import_end(c, genome_size, bloom, bloom_bytes, sketch_size);
}

// Miekki::merge_indexes, Miekki.cpp:901-910: the columns of `other` are appended to ours (the
// reference leaves the Bloom merge as a TODO, :907, and never updates its statistics; here both
// are done).  The result is the index a single in-order build of "our genomes, then theirs" gives:
// fingerprints and statistics do not depend on the Bloom table, and a table byte belongs to the
// smallest (genome id, bucket, probe) that maps to it (sketch.cu: bloom_commit_kernel), i.e. to
// us whenever we have set it -- the same fold as across GPU shards.
int mk_index_merge(mk_ctx* c, mk_ctx* other) {
    if (!c || !other) return fail(c, MK_ERR_ARG, "mk_index_merge: NULL argument");
    if (c == other) return fail(c, MK_ERR_ARG, "mk_index_merge: an index cannot be merged into itself");
    // both contexts stay locked; address order avoids a deadlock between crossing merges
    std::unique_lock<std::mutex> l1(c < other ? c->mu : other->mu), l2(c < other ? other->mu : c->mu);
    if (c->k != other->k || c->h != other->h || c->b != other->b || c->nbm != other->nbm || c->nbmant != other->nbmant)
        return fail(c, MK_ERR_ARG, "mk_index_merge: the two indexes were built with different -k / -h / -f / -b");
    if (c->import_open || other->import_open) return fail(c, MK_ERR_STATE, "mk_index_merge: an import is in progress");
    // everything queued on the other index must have landed before its memory is read from here
    cudaSetDevice(other->device);
    if (cudaStreamSynchronize(other->stream) != cudaSuccess)
        return fail(c, MK_ERR_CUDA, "mk_index_merge: the other context's stream failed");
    cudaSetDevice(c->device);
    const uint32_t m = other->n;
    if ((uint64_t)c->n + m > 0xFFFFFFFFull) return fail(c, MK_ERR_ARG, "mk_index_merge: more than 2^32 genomes");
    if (m) {
        TRY(ensure_capacity(c, c->n + m));
        // rows travel in slabs (cudaMemcpyDefault: the two contexts may sit on different GPUs)
        const uint64_t slab = std::max<uint64_t>(1, (256ull << 20) / other->stride);
        TRY(reserve(c, c->misc, std::min<uint64_t>(slab, c->B) * other->stride));
        for (uint64_t r0 = 0; r0 < c->B; r0 += slab) {
            const uint64_t nr = std::min<uint64_t>(slab, c->B - r0);
            CU(cudaMemcpyAsync(c->misc.p, other->rows + r0 * other->stride, nr * other->stride, cudaMemcpyDefault,
                               c->stream));
            launch_merge_planes(c->rows, c->stride, c->n, static_cast<const uint8_t*>(c->misc.p), other->stride, m,
                                r0, nr, c->stream);
            c->stats.kernel_launches += 1;
        }
        CU(cudaMemcpyAsync(c->d_sketch_size + c->n, other->d_sketch_size, (size_t)m * 4, cudaMemcpyDefault, c->stream));
        CU(cudaMemcpyAsync(c->d_genome_size + c->n, other->d_genome_size, (size_t)m * 8, cudaMemcpyDefault, c->stream));
        CU(cudaMemcpyAsync(c->d_ratio + c->n, other->d_ratio, (size_t)m * 4, cudaMemcpyDefault, c->stream));
    }
    TRY(reserve(c, c->misc, c->window));
    CU(cudaMemcpyAsync(c->misc.p, other->bloom, c->window, cudaMemcpyDefault, c->stream));
    launch_bloom_merge(c->bloom, static_cast<uint8_t*>(c->misc.p), c->window, c->stream);
    c->stats.kernel_launches += 1;
    CU(cudaGetLastError());
    c->n += m;           // the host mirrors of the statistics are completed on demand (host_stats)
    return sync(c);
}

uint64_t mk_bloom_window(const mk_ctx* c) { return c ? c->window : 0; }

// A canonical k-mer min(S, RC) is < M = 4^k - 4^(ceil(k/2)-1): a digit 3 of S comes from a 'T'
// (or 't' in the prefix encoder), whose reverse-strand digit is 0, so S >= 4^k - 4^j forces the
// low k-j digits of RC to 0, i.e. RC <= 4^k - 4^(k-j), and both can hold only for j >= k/2.
// Table bytes from ((M - 1 + 1023) >> (b + 3)) + 1 on are therefore never probed; the build's
// "page is saturated" test ignores them.  Pure arithmetic (no device needed):
// tests/test_abi_cpu.py checks the bound against every k-mer the oracle can produce for small k.
uint64_t mk_bloom_reach(uint32_t k, uint32_t bloom_log2) {
    if (k < 1 || k > 31 || bloom_log2 > 60) return 0;
    const uint64_t M = (1ull << (2 * k)) - (1ull << (2 * ((k + 1) / 2 - 1)));
    return (((M - 1 + 1023) >> bloom_log2) >> 3) + 1;
}

int mk_bloom_get(mk_ctx* c, uint8_t* dst, uint64_t n) {
    if (!c || !dst) return MK_ERR_ARG;
    Guard g(c);
    c->prep.serial = 0;        // lists sketched ahead saw the old index
    if (n > c->window) return fail(c, MK_ERR_ARG, "mk_bloom_get: n exceeds the window");
    CU(cudaMemcpyAsync(dst, c->bloom, n, cudaMemcpyDefault, c->stream));
    c->stats.d2h_bytes += n;
    return sync(c);
}

int mk_bloom_set(mk_ctx* c, const uint8_t* src, uint64_t n) {
    if (!c || !src) return MK_ERR_ARG;
    Guard g(c);
    c->prep.serial = 0;        // lists sketched ahead saw the old index
    if (n > c->window) return fail(c, MK_ERR_ARG, "mk_bloom_set: n exceeds the window");
    CU(cudaMemcpyAsync(c->bloom, src, n, cudaMemcpyDefault, c->stream));
    c->stats.h2d_bytes += n;
    return sync(c);
}

int mk_bloom_merge(mk_ctx* c, const uint8_t* src, uint64_t n) {
    if (!c || !src) return MK_ERR_ARG;
    Guard g(c);
    c->prep.serial = 0;        // lists sketched ahead saw the old index
    if (n > c->window || n % 16) return fail(c, MK_ERR_ARG, "mk_bloom_merge: n must be a multiple of 16 within the window");
    TRY(reserve(c, c->misc, n));
    CU(cudaMemcpyAsync(c->misc.p, src, n, cudaMemcpyDefault, c->stream));
    launch_bloom_merge(c->bloom, static_cast<uint8_t*>(c->misc.p), n, c->stream);
    c->stats.kernel_launches += 1;
    c->stats.h2d_bytes += n;
    return sync(c);
}

// ---- query ---------------------------------------------------------------------------

int mk_query_batch(mk_ctx* c, const mk_batch* reads, uint32_t nresults, uint32_t min_score,
                   double min_intersection, mk_hit* hits, uint32_t* nhits) {
    if (!c || !reads) return fail(c, MK_ERR_ARG, "mk_query_batch: NULL argument");
    Guard g(c);
    return query_device(c, reads, nresults, min_score, min_intersection, hits, nhits, false, 1);
}

int mk_query(mk_ctx* c, const char* const* seqs, const uint64_t* lens, uint32_t n, uint32_t nresults,
             uint32_t min_score, double min_intersection, mk_hit* hits, uint32_t* nhits) {
    if (!c || (n && (!seqs || !lens || !hits || !nhits))) return fail(c, MK_ERR_ARG, "mk_query: NULL argument");
    Guard g(c);
    mk_batch* b = nullptr;
    TRY(upload_range(c, seqs, lens, n, &b));
    int r = query_device(c, b, nresults, min_score, min_intersection, hits, nhits, false, 1);
    cudaStreamSynchronize(c->stream);
    batch_release(b);
    return r;
}

int mk_query_chain(mk_ctx* c, const mk_batch* reads, uint32_t nresults, uint32_t min_score,
                   double min_intersection, mk_hit* heap_io, uint32_t* len_io, int finalize) {
    if (!c || !reads) return fail(c, MK_ERR_ARG, "mk_query_chain: NULL argument");
    Guard g(c);
    return query_device(c, reads, nresults, min_score, min_intersection, heap_io, len_io, true, finalize);
}

int mk_scan(mk_ctx* c, const mk_batch* reads) {
    if (!c || !reads) return fail(c, MK_ERR_ARG, "mk_scan: NULL argument");
    Guard g(c);
    TRY(scan_all_async(c, reads, nullptr));
    return sync(c);
}

int mk_scan_async(mk_ctx* c, const mk_batch* reads, int* slot) {
    if (!c || !reads) return fail(c, MK_ERR_ARG, "mk_scan_async: NULL argument");
    Guard g(c);
    return scan_all_async(c, reads, slot);
}

int mk_sketch_async(mk_ctx* c, const mk_batch* reads) {
    if (!c || !reads) return fail(c, MK_ERR_ARG, "mk_sketch_async: NULL argument");
    Guard g(c);
    return sketch_ahead(c, reads);
}

int mk_topk(mk_ctx* c, uint32_t nresults, uint32_t min_score, double min_intersection, mk_hit* heap_io,
            uint32_t* len_io, int chain_in, int finalize) {
    if (!c) return MK_ERR_ARG;
    Guard g(c);
    return topk_slot(c, c->last_slot, 0, UINT32_MAX, nresults, min_score, min_intersection, heap_io, len_io, chain_in != 0, finalize);
}

int mk_topk_slot(mk_ctx* c, int slot, uint32_t nresults, uint32_t min_score, double min_intersection,
                 mk_hit* heap_io, uint32_t* len_io, int chain_in, int finalize) {
    if (!c) return MK_ERR_ARG;
    Guard g(c);
    return topk_slot(c, slot, 0, UINT32_MAX, nresults, min_score, min_intersection, heap_io, len_io, chain_in != 0, finalize);
}

int mk_topk_slot_range(mk_ctx* c, int slot, uint32_t first_read, uint32_t n_reads, uint32_t nresults,
                       uint32_t min_score, double min_intersection, mk_hit* heap_io, uint32_t* len_io, int chain_in,
                       int finalize) {
    if (!c) return MK_ERR_ARG;
    Guard g(c);
    return topk_slot(c, slot, first_read, n_reads, nresults, min_score, min_intersection, heap_io, len_io,
                     chain_in != 0, finalize);
}

int mk_query_counts(mk_ctx* c, const char* const* seqs, const uint64_t* lens, uint32_t n, uint32_t* counts,
                    uint32_t* surviving) {
    if (!c || (n && (!seqs || !lens || !counts))) return fail(c, MK_ERR_ARG, "mk_query_counts: NULL argument");
    Guard g(c);
    if (n == 0) return MK_OK;
    mk_batch* b = nullptr;
    TRY(upload_range(c, seqs, lens, n, &b));
    auto body = [&]() -> int {
        Lists L{};
        TRY(build_lists(c, b, 0, n, &L));
        if (c->n) {
            ScanPlan plan{};
            if (scan_plan(c->n, c->stride, c->sm_count, c->smem_optin, c->scan_spare_sms, &plan) != 0)
                return fail(c, MK_ERR_CUDA, "no scan plan for this index width");
            const uint32_t qb = scan_batch_reads(c, n);
            const uint64_t n_pad = (c->n + 31) / 32 * 32;
            TRY(reserve(c, c->counts, (size_t)qb * n_pad * 4));
            for (uint32_t q0 = 0; q0 < n; q0 += qb) {
                const uint32_t nq = std::min(qb, n - q0);
                TRY(scan_reads(c, L, q0, nq, plan));
                CU(cudaMemcpy2DAsync(counts + (size_t)q0 * c->n, (size_t)c->n * 4, c->counts.p, n_pad * 4,
                                     (size_t)c->n * 4, nq, cudaMemcpyDeviceToHost, c->stream));
                CU(cudaStreamSynchronize(c->stream));
                c->stats.d2h_bytes += (size_t)nq * c->n * 4;
            }
        }
        TRY(account_lists(c, L, n, surviving));
        return sync(c);
    };
    int r = body();
    cudaStreamSynchronize(c->stream);
    batch_release(b);
    return r;
}

int mk_sketch(mk_ctx* c, const char* seq, uint64_t len, uint8_t* fp, uint64_t* anc, uint32_t* active) {
    if (!c || (len && !seq)) return fail(c, MK_ERR_ARG, "mk_sketch: NULL argument");
    Guard g(c);
    mk_batch* b = nullptr;
    const char* seqs[1] = {seq};
    TRY(upload_range(c, seqs, &len, 1, &b));
    auto body = [&]() -> int {
        BatchView v = view_of(b, 0, 1);
        DenseOut d{};
        {
            PhaseTimer t(c, PH_SKETCH);
            TRY(dense_sketch(c, v, false, &d));
        }
        if (fp) CU(cudaMemcpyAsync(fp, c->fp.p, c->B, cudaMemcpyDeviceToHost, c->stream));
        if (anc) CU(cudaMemcpyAsync(anc, c->keys.p, c->B * 8, cudaMemcpyDeviceToHost, c->stream));
        if (active) CU(cudaMemcpyAsync(active, d.d_active, 4, cudaMemcpyDeviceToHost, c->stream));
        c->stats.d2h_bytes += c->B * 9 + 4;
        return sync(c);
    };
    int r = body();
    cudaStreamSynchronize(c->stream);
    batch_release(b);
    return r;
}

// ---- exact mode ----------------------------------------------------------------------

// set B from the records of `gb`, then every read of `rb` against it (both resident in HBM)
static int exact_device(mk_ctx* c, const mk_batch* gb, const mk_batch* rb, uint64_t* nb_inter, uint64_t* nb_union,
                        uint64_t* genome_distinct) {
    unsigned long long *tableB = nullptr, *rtable = nullptr, *d_cnt = nullptr;
    uint64_t* d_toff = nullptr;
    const uint32_t n_records = gb->n, n_reads = rb->n;
    const uint64_t* rec_lens = gb->h_len.data();
    const uint64_t* read_lens = rb->h_len.data();
    auto body = [&]() -> int {
        const uint32_t k = c->k;
        uint64_t wins = 0;
        for (uint32_t i = 0; i < n_records; ++i)
            if (rec_lens[i] >= k) wins += rec_lens[i] - k + 1;
        const uint64_t slotsB = std::max<uint64_t>(16, wins * 2);
        std::vector<uint64_t> toff((size_t)n_reads + 1);
        uint64_t tt = 0;
        for (uint32_t i = 0; i < n_reads; ++i) {
            toff[i] = tt;
            tt += read_lens[i] >= k ? std::max<uint64_t>(4, (read_lens[i] - k + 1) * 2) : 1;
        }
        toff[n_reads] = tt;
        // MIEKKI_EXACT_SORT=1: set B as a sorted array (keys -> radix sort -> binary search) instead of
        // a hash set; same results, kept for the comparison in DESIGN.md section 4 (read per call)
        const char* se = getenv("MIEKKI_EXACT_SORT");
        const bool sorted_b = se && atoi(se) != 0 && wins > 0;
        std::vector<uint64_t> koff((size_t)n_records + 1, 0);
        for (uint32_t i = 0; i < n_records; ++i) koff[i + 1] = koff[i] + (rec_lens[i] >= k ? rec_lens[i] - k + 1 : 0);
        const size_t sort_tmp = sorted_b ? exact_sort_temp_bytes(wins, (int)k) : 0;
        // hash set: 2 slots per window; sorted: keys | sorted keys | koff | cub scratch
        const size_t b_bytes = sorted_b ? wins * 16 + ((size_t)n_records + 1) * 8 + sort_tmp + 256 : slotsB * 8;
        CU(cudaMallocAsync(reinterpret_cast<void**>(&tableB), b_bytes, c->stream));
        CU(cudaMallocAsync(reinterpret_cast<void**>(&rtable), std::max<uint64_t>(1, tt) * 8, c->stream));
        CU(cudaMallocAsync(reinterpret_cast<void**>(&d_cnt), (1 + 2 * (size_t)n_reads) * 8, c->stream));
        CU(cudaMallocAsync(reinterpret_cast<void**>(&d_toff), ((size_t)n_reads + 1) * 8, c->stream));
        PhaseTimer t(c, PH_EXACT);
        if (!sorted_b) launch_fill_u64(tableB, slotsB, ~0ull, c->stream);
        launch_fill_u64(rtable, tt, ~0ull, c->stream);
        CU(cudaMemsetAsync(d_cnt, 0, (1 + 2 * (size_t)n_reads) * 8, c->stream));
        CU(cudaMemcpyAsync(d_toff, toff.data(), ((size_t)n_reads + 1) * 8, cudaMemcpyHostToDevice, c->stream));
        c->stats.kernel_launches += 2;
        const uint32_t YMAX = 32768;
        unsigned long long* sortedB = tableB + wins;
        if (sorted_b) {
            uint64_t* d_koff = reinterpret_cast<uint64_t*>(tableB + 2 * wins);
            void* d_tmp = reinterpret_cast<uint8_t*>(d_koff + n_records + 1) + (256 - ((uintptr_t)(d_koff + n_records + 1) & 255)) % 256;
            CU(cudaMemcpyAsync(d_koff, koff.data(), ((size_t)n_records + 1) * 8, cudaMemcpyHostToDevice, c->stream));
            if (n_records > YMAX) return fail(c, MK_ERR_ARG, "mk_exact (sorted variant): more than 32,768 records");
            BatchView v = view_of(gb, 0, n_records);
            launch_exact_sorted_build(v.chars, v.d_coff, v.d_len, d_koff, n_records, v.max_len, wins, (int)k, tableB,
                                      sortedB, d_tmp, sort_tmp, d_cnt, c->stream);
            CU(cudaStreamSynchronize(c->stream));     // koff is a local vector
            c->stats.kernel_launches += 3;
        } else {
            for (uint32_t f = 0; f < n_records; f += YMAX) {
                const uint32_t m = std::min(YMAX, n_records - f);
                BatchView v = view_of(gb, f, m);
                launch_exact_insert(v.chars, v.d_coff, v.d_len, m, v.max_len, (int)k, tableB, slotsB, d_cnt, c->stream);
                c->stats.kernel_launches += 1;
            }
        }
        unsigned long long* d_inter = d_cnt + 1;
        unsigned long long* d_dist = d_cnt + 1 + n_reads;
        for (uint32_t f = 0; f < n_reads; f += YMAX) {
            const uint32_t m = std::min(YMAX, n_reads - f);
            BatchView v = view_of(rb, f, m);
            if (sorted_b)
                launch_exact_reads_sorted(v.chars, v.d_coff, v.d_len, m, v.max_len, (int)k, rtable, d_toff + f, sortedB,
                                          wins, d_inter + f, d_dist + f, c->stream);
            else
                launch_exact_reads(v.chars, v.d_coff, v.d_len, m, v.max_len, (int)k, rtable, d_toff + f, tableB, slotsB,
                                   d_inter + f, d_dist + f, c->stream);
            c->stats.kernel_launches += 1;
        }
        CU(cudaGetLastError());
        std::vector<unsigned long long> cnt(1 + 2 * (size_t)n_reads);
        CU(cudaMemcpyAsync(cnt.data(), d_cnt, cnt.size() * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->stats.d2h_bytes += cnt.size() * 8;
        if (genome_distinct) *genome_distinct = cnt[0];
        for (uint32_t i = 0; i < n_reads; ++i) {
            nb_inter[i] = cnt[1 + i];
            nb_union[i] = cnt[0] + (cnt[1 + n_reads + i] - cnt[1 + i]);   // |B| + |A \ B|
        }
        return MK_OK;
    };
    int r = body();
    cudaStreamSynchronize(c->stream);
    resolve_events(c);
    if (tableB) cudaFreeAsync(tableB, c->stream);
    if (rtable) cudaFreeAsync(rtable, c->stream);
    if (d_cnt) cudaFreeAsync(d_cnt, c->stream);
    if (d_toff) cudaFreeAsync(d_toff, c->stream);
    return r;
}

// Several genomes in one call (mk_exact_many): one hash table (sized for the largest genome) is
// reused, nothing waits for the host between genomes, the counters come back in one copy -- and the
// genomes travel in sub-batches of about 32 MB on the copy stream, so that the upload of one
// sub-batch runs beside the kernels of the previous one.
int mk_exact_many(mk_ctx* c, uint32_t n_genomes, const char* const* records, const uint64_t* rec_lens,
                  const uint32_t* rec_count, const char* const* reads, const uint64_t* read_lens,
                  const uint32_t* read_count, uint64_t* nb_inter, uint64_t* nb_union, uint64_t* genome_distinct) {
    if (!c || (n_genomes && (!rec_count || !read_count))) return fail(c, MK_ERR_ARG, "mk_exact_many: NULL argument");
    uint64_t n_records = 0, n_reads64 = 0;
    for (uint32_t g = 0; g < n_genomes; ++g) {
        n_records += rec_count[g];
        n_reads64 += read_count[g];
    }
    if (n_records > 0xFFFFFFFFull || n_reads64 > 0xFFFFFFFFull) return fail(c, MK_ERR_ARG, "mk_exact_many: too many sequences");
    if ((n_records && (!records || !rec_lens)) || (n_reads64 && (!reads || !read_lens || !nb_inter || !nb_union)))
        return fail(c, MK_ERR_ARG, "mk_exact_many: NULL argument");
    Guard guard(c);
    const uint32_t n_reads = (uint32_t)n_reads64, k = c->k;
    unsigned long long *tableB = nullptr, *rtable = nullptr, *d_cnt = nullptr;
    uint64_t* d_toff = nullptr;
    mk_batch* rb = nullptr;
    std::vector<mk_batch*> subs;
    auto body = [&]() -> int {
        TRY(upload_range(c, reads, read_lens, n_reads, &rb));
        std::vector<uint64_t> wins(n_genomes, 0);
        uint64_t max_wins = 0;
        for (uint32_t g = 0, r0 = 0; g < n_genomes; r0 += rec_count[g], ++g) {
            for (uint32_t i = 0; i < rec_count[g]; ++i)
                if (rec_lens[r0 + i] >= k) wins[g] += rec_lens[r0 + i] - k + 1;
            max_wins = std::max(max_wins, wins[g]);
        }
        std::vector<uint64_t> toff((size_t)n_reads + 1);
        uint64_t tt = 0;
        for (uint32_t i = 0; i < n_reads; ++i) {
            toff[i] = tt;
            tt += read_lens[i] >= k ? std::max<uint64_t>(4, (read_lens[i] - k + 1) * 2) : 1;
        }
        toff[n_reads] = tt;
        const size_t n_cnt = (size_t)n_genomes + 2 * (size_t)n_reads;      // |B| per genome | inter | |A| per read
        CU(cudaMallocAsync(reinterpret_cast<void**>(&tableB), std::max<uint64_t>(16, max_wins * 2) * 8, c->stream));
        CU(cudaMallocAsync(reinterpret_cast<void**>(&rtable), std::max<uint64_t>(1, tt) * 8, c->stream));
        CU(cudaMallocAsync(reinterpret_cast<void**>(&d_cnt), std::max<size_t>(1, n_cnt) * 8, c->stream));
        CU(cudaMallocAsync(reinterpret_cast<void**>(&d_toff), ((size_t)n_reads + 1) * 8, c->stream));
        PhaseTimer t(c, PH_EXACT);
        launch_fill_u64(rtable, tt, ~0ull, c->stream);
        CU(cudaMemsetAsync(d_cnt, 0, std::max<size_t>(1, n_cnt) * 8, c->stream));
        CU(cudaMemcpyAsync(d_toff, toff.data(), ((size_t)n_reads + 1) * 8, cudaMemcpyHostToDevice, c->stream));
        c->stats.kernel_launches += 1;
        unsigned long long* d_inter = d_cnt + n_genomes;
        unsigned long long* d_dist = d_inter + n_reads;
        const uint32_t YMAX = 32768;
        uint32_t g0 = 0, r0 = 0, q0 = 0;
        while (g0 < n_genomes) {
            // the next sub-batch of genomes: about 32 MB of records
            uint32_t g1 = g0, nrec = 0;
            uint64_t bytes = 0;
            while (g1 < n_genomes && (g1 == g0 || bytes < (32ull << 20))) {
                for (uint32_t i = 0; i < rec_count[g1]; ++i) bytes += rec_lens[r0 + nrec + i];
                nrec += rec_count[g1];
                ++g1;
            }
            mk_batch* gb = nullptr;
            // returns when the copy has landed; the kernels of the previous sub-batch keep running
            TRY(upload_range(c, records + r0, rec_lens + r0, nrec, &gb, c->copy_stream));
            subs.push_back(gb);
            for (uint32_t g = g0, rr = 0; g < g1; rr += rec_count[g], q0 += read_count[g], ++g) {
                const uint64_t slotsB = std::max<uint64_t>(16, wins[g] * 2);
                launch_fill_u64(tableB, slotsB, ~0ull, c->stream);
                c->stats.kernel_launches += 1;
                for (uint32_t f = 0; f < rec_count[g]; f += YMAX) {
                    const uint32_t m = std::min(YMAX, rec_count[g] - f);
                    BatchView v = view_of(gb, rr + f, m);
                    launch_exact_insert(v.chars, v.d_coff, v.d_len, m, v.max_len, (int)k, tableB, slotsB, d_cnt + g, c->stream);
                    c->stats.kernel_launches += 1;
                }
                for (uint32_t f = 0; f < read_count[g]; f += YMAX) {
                    const uint32_t m = std::min(YMAX, read_count[g] - f);
                    BatchView v = view_of(rb, q0 + f, m);
                    launch_exact_reads(v.chars, v.d_coff, v.d_len, m, v.max_len, (int)k, rtable, d_toff + q0 + f, tableB,
                                       slotsB, d_inter + q0 + f, d_dist + q0 + f, c->stream);
                    c->stats.kernel_launches += 1;
                }
            }
            r0 += nrec;
            g0 = g1;
        }
        CU(cudaGetLastError());
        std::vector<unsigned long long> cnt(std::max<size_t>(1, n_cnt));
        CU(cudaMemcpyAsync(cnt.data(), d_cnt, cnt.size() * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->stats.d2h_bytes += cnt.size() * 8;
        for (uint32_t g = 0, qq = 0; g < n_genomes; qq += read_count[g], ++g) {
            if (genome_distinct) genome_distinct[g] = cnt[g];
            for (uint32_t i = qq; i < qq + read_count[g]; ++i) {
                nb_inter[i] = cnt[n_genomes + i];
                nb_union[i] = cnt[g] + (cnt[n_genomes + n_reads + i] - cnt[n_genomes + i]);   // |B| + |A \ B|
            }
        }
        return MK_OK;
    };
    int r = body();
    cudaStreamSynchronize(c->stream);
    resolve_events(c);
    if (tableB) cudaFreeAsync(tableB, c->stream);
    if (rtable) cudaFreeAsync(rtable, c->stream);
    if (d_cnt) cudaFreeAsync(d_cnt, c->stream);
    if (d_toff) cudaFreeAsync(d_toff, c->stream);
    for (mk_batch* gb : subs) batch_release(gb);
    batch_release(rb);
    return r;
}

int mk_exact(mk_ctx* c, const char* const* records, const uint64_t* rec_lens, uint32_t n_records,
             const char* const* reads, const uint64_t* read_lens, uint32_t n_reads, uint64_t* nb_inter,
             uint64_t* nb_union, uint64_t* genome_distinct) {
    if (!c || (n_records && (!records || !rec_lens)) || (n_reads && (!reads || !read_lens || !nb_inter || !nb_union)))
        return fail(c, MK_ERR_ARG, "mk_exact: NULL argument");
    Guard g(c);
    mk_batch *gb = nullptr, *rb = nullptr;
    int r = upload_range(c, records, rec_lens, n_records, &gb);
    if (r == MK_OK) r = upload_range(c, reads, read_lens, n_reads, &rb);
    if (r == MK_OK) r = exact_device(c, gb, rb, nb_inter, nb_union, genome_distinct);
    cudaStreamSynchronize(c->stream);
    batch_release(gb);
    batch_release(rb);
    return r;
}

int mk_exact_batch(mk_ctx* c, const mk_batch* records, const mk_batch* reads, uint64_t* nb_inter,
                   uint64_t* nb_union, uint64_t* genome_distinct) {
    if (!c || !records || !reads || (reads->n && (!nb_inter || !nb_union)))
        return fail(c, MK_ERR_ARG, "mk_exact_batch: NULL argument");
    Guard g(c);
    return exact_device(c, records, reads, nb_inter, nb_union, genome_distinct);
}

// ---- measurement ---------------------------------------------------------------------

// folds the device-side work counters into the host statistics
static int fold_device_stats(mk_ctx* c) {
    CU(cudaStreamSynchronize(c->aux_stream));
    TRY(sync(c));
    unsigned long long st[2] = {0, 0};
    CU(cudaMemcpy(st, c->d_stat, sizeof(st), cudaMemcpyDeviceToHost));
    CU(cudaMemset(c->d_stat, 0, sizeof(st)));
    c->stats.scan_rows += st[0];
    c->stats.scan_row_bytes += st[1];
    c->stats.h2d_bytes += c->h2d_meta.exchange(0, std::memory_order_relaxed);
    return MK_OK;
}

int mk_stats_get(mk_ctx* c, mk_stats* out) {
    if (!c || !out) return MK_ERR_ARG;
    Guard g(c);
    TRY(fold_device_stats(c));
    *out = c->stats;
    return MK_OK;
}

int mk_stats_reset(mk_ctx* c) {
    if (!c) return MK_ERR_ARG;
    Guard g(c);
    TRY(fold_device_stats(c));
    c->stats = mk_stats{};
    return MK_OK;
}

int mk_sync(mk_ctx* c) {
    if (!c) return MK_ERR_ARG;
    Guard g(c);
    return sync(c);
}

}  // extern "C"
