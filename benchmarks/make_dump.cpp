// make_dump -- bench tooling: materialises the index of a synthetic configuration as a dump
// file the UNMODIFIED reference binary loads with -i (plain bytes: zstr reads uncompressed
// files as they are, zstr.hpp:154-167; layout Miekki.cpp:651-676).
//
// `bench.py --impl reference` runs this as a separate step and then times only the reference
// binary (oracle/_ref/Miekki) in a process that never maps libmiekki_b200.so.  Building the
// 10,000-genome index of BASELINE config 2 with the reference itself takes about an hour on the
// box's host cores; the GPU build is bit-identical to a reference `-t 1` build (every row,
// statistic and Bloom byte: tests/test_gpu_parity.py, tests/test_gpu_config1.py), so the index is
// input data here, like the reads, not part of what is timed.
//
//   make_dump <k> <h> <b> <threshold> <n_genomes> <genome_len> <seed> <first_genome> <out>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "miekki_b200.h"

static void die(mk_ctx* c, const char* what) {
    fprintf(stderr, "make_dump: %s: %s\n", what, mk_last_error(c));
    exit(1);
}
static void put(FILE* f, const void* p, size_t n) {
    if (n && fwrite(p, 1, n, f) != n) {
        fprintf(stderr, "make_dump: short write\n");
        exit(1);
    }
}

int main(int argc, char** argv) {
    if (argc != 10) {
        fprintf(stderr, "usage: make_dump k h b threshold n_genomes genome_len seed first_genome out\n");
        return 2;
    }
    const uint32_t k = (uint32_t)atoi(argv[1]), h = (uint32_t)atoi(argv[2]), b = (uint32_t)atoi(argv[3]);
    const uint32_t threshold = (uint32_t)atoi(argv[4]), n = (uint32_t)atoi(argv[5]);
    const uint64_t len = strtoull(argv[6], nullptr, 10), seed = strtoull(argv[7], nullptr, 0);
    const uint32_t first = (uint32_t)atoi(argv[8]);
    const char* out = argv[9];
    const auto t0 = std::chrono::steady_clock::now();
    mk_ctx* c = nullptr;
    if (mk_create(k, h, 8, 5, b, threshold, 0, &c) != MK_OK) die(nullptr, "mk_create");
    if (mk_index_reserve(c, n) != MK_OK) die(c, "mk_index_reserve");
    for (uint32_t g0 = 0; g0 < n; g0 += 128) {
        const uint32_t m = n - g0 < 128 ? n - g0 : 128;
        mk_batch* bt = nullptr;
        if (mk_batch_synth(c, seed, first + g0, m, len, &bt) != MK_OK) die(c, "mk_batch_synth");
        if (mk_index_add_batch(c, bt) != MK_OK) die(c, "mk_index_add_batch");
        mk_batch_free(c, bt);
    }
    const auto t1 = std::chrono::steady_clock::now();
    FILE* f = fopen(out, "wb");
    if (!f) {
        fprintf(stderr, "make_dump: cannot open %s\n", out);
        return 1;
    }
    const uint64_t B = 1ull << h, bloom_bits = 1ull << b;
    const uint32_t nbm = 8, nbmant = 5;
    const uint8_t flags[2] = {0, 0}, compressed = 0;
    put(f, &k, 4); put(f, &h, 4); put(f, &nbm, 4); put(f, &nbmant, 4); put(f, &n, 4); put(f, &b, 4);
    put(f, &bloom_bits, 8); put(f, flags, 2); put(f, &threshold, 4); put(f, &compressed, 1);
    const uint64_t slab = (256ull << 20) / (n ? n : 1) ? (256ull << 20) / (n ? n : 1) : 1;
    std::vector<uint8_t> buf;
    for (uint64_t r0 = 0; r0 < B && n; r0 += slab) {
        const uint64_t nr = B - r0 < slab ? B - r0 : slab;
        buf.resize(nr * n);
        if (mk_index_export_rows(c, r0, nr, buf.data(), n) != MK_OK) die(c, "mk_index_export_rows");
        put(f, buf.data(), buf.size());
    }
    std::vector<uint64_t> gs(n);
    std::vector<uint32_t> ss(n);
    const uint64_t window = mk_bloom_window(c) < bloom_bits / 8 ? mk_bloom_window(c) : bloom_bits / 8;
    std::vector<uint8_t> bloom(window);
    if (mk_index_export(c, nullptr, gs.data(), bloom.data(), bloom.size(), ss.data()) != MK_OK) die(c, "mk_index_export");
    put(f, gs.data(), gs.size() * 8);
    put(f, bloom.data(), bloom.size());
    std::vector<uint8_t> zeros(16u << 20, 0);
    for (uint64_t left = bloom_bits / 8 - window; left;) {
        const size_t m = left < zeros.size() ? (size_t)left : zeros.size();
        put(f, zeros.data(), m);
        left -= m;
    }
    put(f, ss.data(), ss.size() * 4);
    if (fclose(f) != 0) {
        fprintf(stderr, "make_dump: close failed\n");
        return 1;
    }
    mk_destroy(c);
    const auto t2 = std::chrono::steady_clock::now();
    printf("{\"genomes\": %u, \"build_s\": %.3f, \"write_s\": %.3f}\n", n, std::chrono::duration<double>(t1 - t0).count(),
           std::chrono::duration<double>(t2 - t1).count());
    return 0;
}
