// miekki -- command line of the B200-native sketch-and-query path.
//
// Drop-in for the reference's main.cpp:125-238 on the flags -l -a -A -o -d -i -h -k -s -f -b
// -e -t: same getopt string, same defaults (h=17, t=8, k=31, b=33, f=3, s=200, out.txt), same
// hit-line format (Miekki.cpp:438-445), exact-line format (Miekki.cpp:853) and gz index dump
// (Miekki.cpp:649-719).  The host keeps file parsing (plain/gz FASTA) and text formatting;
// sketching, scoring, filtering and the exact intersection run on the GPU through the C ABI
// in include/miekki_b200.h.  Genome ids are list order (the reference's `-t 1` order) for any
// -t; -t only sets the number of host parser threads.
//
// --gpus N shards the genomes (matrix columns) over N GPUs of the box in contiguous ascending
// id ranges: every GPU builds and scans its own columns, the Bloom table is
// folded "lowest shard wins" after the build and the bounded heap is chained through the
// shards in id order, so every output byte is the same as with one GPU.
#include <getopt.h>
#include <omp.h>
#include <sys/stat.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <future>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "fasta.hpp"
#include "miekki_b200.h"

using namespace std;

namespace {

// MIEKKI_TIMING=1: phase timings on stderr (stdout stays what the reference prints)
struct PhaseClock {
    const char* what;
    chrono::steady_clock::time_point t0 = chrono::steady_clock::now();
    static bool on() {
        static const bool v = getenv("MIEKKI_TIMING") != nullptr;
        return v;
    }
    explicit PhaseClock(const char* w) : what(w) {}
    ~PhaseClock() {
        if (!on()) return;
        const double s = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
        #pragma omp critical(msg)
        cerr << "[timing] " << what << " " << s << " s" << endl;
    }
};

bool exists_test(const string& name) {
    struct stat buffer;
    return stat(name.c_str(), &buffer) == 0;
}

// Decimal with a comma between groups of three digits, as the reference's banners print sizes
// (main.cpp:186, Miekki.cpp:585).
string int_to_string(uint64_t n) {
    string digits = to_string(n), out;
    const size_t lead = digits.size() % 3;
    for (size_t i = 0; i < digits.size(); ++i) {
        if (i && i % 3 == lead) out += ',';
        out += digits[i];
    }
    return out;
}

// Readers throw on unreadable / corrupt input (fasta.hpp).  An exception must not cross an OpenMP
// region, so loop bodies run through this: the first message is kept and reported by the caller.
struct LoopError {
    string first;
    template <class F>
    void run(F&& body) {
        try {
            body();
        } catch (const exception& e) {
            #pragma omp critical(loop_error)
            if (first.empty()) first = e.what();
        }
    }
    void rethrow() const {
        if (!first.empty()) throw runtime_error(first);
    }
};

// 64-bit FNV-1a, chained
uint64_t fnv1a(const void* p, size_t n, uint64_t h = 0xCBF29CE484222325ull) {
    const unsigned char* s = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; ++i) h = (h ^ s[i]) * 0x100000001B3ull;
    return h;
}
// What binds a `<dump>.names` side-car to its dump: the header fields and both statistics arrays.
string names_token(uint32_t k, uint32_t h, uint32_t b, uint32_t threshold, const vector<uint32_t>& ss,
                   const vector<uint64_t>& gs) {
    const uint32_t head[5] = {k, h, b, threshold, (uint32_t)ss.size()};
    uint64_t t = fnv1a(head, sizeof(head));
    t = fnv1a(ss.data(), ss.size() * 4, t);
    t = fnv1a(gs.data(), gs.size() * 8, t);
    char buf[32];
    snprintf(buf, sizeof(buf), "%016llx", (unsigned long long)t);
    return buf;
}
const char* const kNamesMagic = "#miekki-names v1 ";

// One table drives both the usage text and the option parser.  Letters, arities and defaults
// are the reference's (getopt string "i:l:a:h:t:f:k:s:b:o:ed:A:", main.cpp:130-137); note that -h
// takes a value (it is not "help") and that the real defaults are h=17, t=8, s=200 (quirk G11).
struct OptionDoc {
    const char* section;
    const char* flag;
    const char* text;
};
const OptionDoc kOptions[] = {
    {"Input", "-i <dump>", "load an index written by -d (its k, h, b, s override the flags); repeat -i to merge "
                           "several dumps, ids in the order given (then -d writes the merged index)"},
    {"Input", "-l <list>", "build an index from a file of genome FASTA paths (plain or gz), one per line"},
    {"Input", "-a <fasta>", "query every 2-line record (header, sequence) of a FASTA file"},
    {"Input", "-A <list>", "query whole files: each listed FASTA is one query"},
    {"Output", "-o <file>", "result file (out.txt)"},
    {"Output", "-d <file>", "write the index as a gz dump for later -i"},
    {"Sketch", "-h <n>", "2^n fingerprints per genome (17)"},
    {"Sketch", "-k <n>", "k-mer size, at most 31 (31)"},
    {"Sketch", "-s <x>", "smallest estimated intersection worth reporting (200)"},
    {"Sketch", "-t <n>", "host threads for parsing and formatting (8)"},
    {"Advanced", "-f <n>", "fingerprint mantissa bits; only 3 (8-bit fingerprints) is implemented"},
    {"Advanced", "-b <n>", "log2 of the Bloom filter size in bits, at least 32 (33)"},
    {"Advanced", "-e", "exact mode: recompute the true k-mer intersection of every hit"},
    {"B200", "--gpus <n>", "shard the genomes over n GPUs of this box (1)"},
    {"B200", "--device <n>", "first CUDA device ordinal (0)"},
    {"B200", "--devices a,b,..", "explicit device ordinal per shard"},
};

void help() {
    cout << "This is a help message" << endl;       // first line as in main.cpp:103
    const char* section = "";
    for (const OptionDoc& o : kOptions) {
        if (string(section) != o.section) {
            section = o.section;
            cout << "\n" << section << endl;
        }
        cout << "  " << o.flag << "  " << o.text << endl;
    }
}

[[noreturn]] void die(mk_ctx* ctx, const char* what) {
    cerr << "miekki: " << what << ": " << mk_last_error(ctx) << endl;
    exit(1);
}

struct Index {
    vector<mk_ctx*> shard;               // one context per GPU; ids ascend with the shard number
    vector<uint32_t> first;              // first genome id of each shard
    uint32_t k = 31, h = 17, nbm = 8, nbmant = 5, b = 33, threshold = 200;
    bool compressed_flag = false;        // header byte 38
    vector<string> file_names;           // Miekki.h:58; not part of the dump (quirk G4)
    ofstream* out = nullptr;
    int threads = 8, device0 = 0, gpus = 1;
    vector<int> devices;                 // --devices a,b,...: explicit ordinals (repeats allowed)

    uint32_t size() const {
        uint32_t n = 0, t = 0;
        for (mk_ctx* c : shard) { mk_index_size(c, &t); n += t; }
        return n;
    }
    void create_shards() {
        shard.assign((size_t)gpus, nullptr);
        for (int r = 0; r < gpus; ++r)
            if (mk_create(k, h, nbm, nbmant, b, threshold, devices.empty() ? device0 + r : devices[(size_t)r],
                          &shard[(size_t)r]) != MK_OK) {
                cerr << "miekki: " << mk_last_error(nullptr) << endl;
                exit(1);
            }
    }
    // ids of the shards + the global Bloom table (byte-wise: the lowest shard's non-zero byte wins)
    void seal_shards() {
        first.assign(shard.size(), 0);
        uint32_t acc = 0, t = 0;
        for (size_t r = 0; r < shard.size(); ++r) {
            first[r] = acc;
            mk_set_shard(shard[r], acc);
            mk_index_size(shard[r], &t);
            acc += t;
        }
        if (shard.size() < 2) return;
        const uint64_t w = mk_bloom_window(shard[0]);
        vector<uint8_t> merged(w), other(w);
        if (mk_bloom_get(shard[0], merged.data(), w) != MK_OK) die(shard[0], "mk_bloom_get");
        for (size_t r = 1; r < shard.size(); ++r) {
            if (mk_bloom_get(shard[r], other.data(), w) != MK_OK) die(shard[r], "mk_bloom_get");
            #pragma omp parallel for num_threads(threads) schedule(static)
            for (uint64_t i = 0; i < w; ++i)
                if (merged[i] == 0) merged[i] = other[i];
        }
        for (mk_ctx* c : shard)
            if (mk_bloom_set(c, merged.data(), w) != MK_OK) die(c, "mk_bloom_set");
    }
};

vector<string> read_list(const string& list) {
    vector<string> names;
    mkcli::LineReader in(list);
    string name;
    while (!in.eof()) {
        in.getline(name);
        if (name.size() > 3) names.push_back(name);              // Miekki.cpp:555, :606
    }
    return names;
}

// ---- build: Miekki::index_file_of_file, Miekki.cpp:540-588 ----------------------------------
// files [lo, hi) of the list go to one shard, in list order
void build_shard(Index& ix, mk_ctx* ctx, const vector<string>& names, size_t lo, size_t hi, int threads,
                 vector<string>& kept) {
    // Files are parsed in waves by a team of host threads and inserted in list order, so ids are
    // deterministic; wave i+1 is parsed (inflate, line joins) while wave i is on the GPU.
    // (128+ files per wave: mk_index_add then sees several 64-genome slices and overlaps the host
    // packing of one with the sketching of the previous one)
    const size_t wave = max<size_t>(128, 8 * (size_t)threads);
    if (hi > lo) mk_index_reserve(ctx, (uint32_t)(hi - lo));     // the matrix is sized once, not regrown
    struct Wave {
        vector<string> seqs;
        vector<char> ok;
    };
    auto parse = [&](size_t w0) {
        PhaseClock pc("parse wave");
        Wave w;
        const size_t m = min(wave, hi - w0);
        w.seqs.resize(m);
        w.ok.assign(m, 0);
        LoopError err;
        #pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
        for (size_t i = 0; i < m; ++i) {
            const string& fn = names[w0 + i];
            if (!exists_test(fn)) {
                #pragma omp critical(msg)
                cout << "Missed file: " << fn << endl;           // :557
                continue;
            }
            err.run([&] {
                w.seqs[i] = mkcli::read_genome_concat(fn);
                w.ok[i] = w.seqs[i].size() >= ix.k;              // :569
            });
        }
        err.rethrow();
        return w;
    };
    future<Wave> next;
    if (lo < hi) next = async(launch::async, parse, lo);
    for (size_t w0 = lo; w0 < hi; w0 += wave) {
        Wave cur = next.get();
        if (w0 + wave < hi) next = async(launch::async, parse, w0 + wave);
        const size_t m = cur.seqs.size();
        vector<string>& seqs = cur.seqs;
        vector<char>& ok = cur.ok;
        vector<const char*> ptr;
        vector<uint64_t> len;
        for (size_t i = 0; i < m; ++i) {
            if (!ok[i]) continue;
            ptr.push_back(seqs[i].data());
            len.push_back(seqs[i].size());
            kept.push_back(names[w0 + i]);                       // :303
            #pragma omp critical(msg)
            cout << "-" << flush;                                // :575
        }
        PhaseClock pc("mk_index_add wave");
        if (!ptr.empty() && mk_index_add(ctx, ptr.data(), len.data(), (uint32_t)ptr.size()) != MK_OK)
            die(ctx, "mk_index_add");
    }
}

void index_file_of_file(Index& ix, const string& list) {
    if (!exists_test(list)) {
        cout << "Missed file of file: " << list << endl;
        return;
    }
    const vector<string> names = read_list(list);
    const size_t R = ix.shard.size();
    vector<vector<string>> kept(R);
    if (R == 1) {
        build_shard(ix, ix.shard[0], names, 0, names.size(), ix.threads, kept[0]);
    } else {
        // contiguous slices of the list per GPU, built concurrently; nested OpenMP teams parse
        omp_set_max_active_levels(2);
        const int per = max(1, ix.threads / (int)R);
        LoopError err;
        #pragma omp parallel for num_threads((int)R) schedule(static, 1)
        for (size_t r = 0; r < R; ++r)
            err.run([&] {
                build_shard(ix, ix.shard[r], names, r * names.size() / R, (r + 1) * names.size() / R, per, kept[r]);
            });
        err.rethrow();
    }
    for (auto& v : kept) ix.file_names.insert(ix.file_names.end(), v.begin(), v.end());
    ix.seal_shards();
    cout << endl;
    cout << "Reference indexed: " << ix.size() << endl;          // :583
    cout << "BF size:" << int_to_string(1ull << ix.b) << endl;   // :585
}

// ---- dump / load: Miekki.cpp:649-719, SURVEY.md Appendix C -----------------------------------
// Rows travel in slabs of about SLAB_BYTES so that neither side ever holds the whole matrix in
// host memory (10.5 GB at -h 20 with 10,000 genomes); slab i+1 is exported / parsed while slab
// i is deflated / converted.
const uint64_t SLAB_BYTES = [] {
    const char* e = getenv("MIEKKI_DUMP_SLAB_BYTES");        // tests shrink it to force many slabs
    return e ? max<uint64_t>(1, strtoull(e, nullptr, 10)) : 256ull << 20;
}();

void dump_disk(Index& ix, const string& path) {
    const uint32_t n = ix.size();
    const uint64_t B = 1ull << ix.h;
    const uint64_t bloom_bits = 1ull << ix.b;
    mkcli::ParallelGzWriter w(path, ix.threads);   // gzip level 1 like zstr::ofstream, multi-member
    const uint8_t jaccard_estimation = 0;       // uninitialised in the reference (quirk G7)
    const uint8_t containment_estimation = 0;
    const uint8_t compressed = ix.compressed_flag ? 1 : 0;
    w.write(&ix.k, 4);
    w.write(&ix.h, 4);
    w.write(&ix.nbm, 4);
    w.write(&ix.nbmant, 4);
    w.write(&n, 4);
    w.write(&ix.b, 4);
    w.write(&bloom_bits, 8);
    w.write(&jaccard_estimation, 1);
    w.write(&containment_estimation, 1);
    w.write(&ix.threshold, 4);
    w.write(&compressed, 1);
    if (n) {
        // shard r owns columns [first[r], first[r] + n_r) of every row of the slab
        const uint64_t slab = max<uint64_t>(1, min<uint64_t>(B, SLAB_BYTES / n));
        vector<uint8_t> buf[2];
        auto fetch = [&](uint64_t r0, vector<uint8_t>& dst) {
            const uint64_t nr = min(slab, B - r0);
            dst.resize(nr * n);
            #pragma omp parallel for num_threads((int)ix.shard.size()) schedule(static, 1)
            for (size_t r = 0; r < ix.shard.size(); ++r)
                if (mk_index_export_rows(ix.shard[r], r0, nr, dst.data() + ix.first[r], n) != MK_OK)
                    die(ix.shard[r], "mk_index_export_rows");
        };
        future<void> next = async(launch::async, fetch, (uint64_t)0, ref(buf[0]));
        int cur = 0;
        for (uint64_t r0 = 0; r0 < B; r0 += slab, cur ^= 1) {
            next.get();
            if (r0 + slab < B) next = async(launch::async, fetch, r0 + slab, ref(buf[cur ^ 1]));
            w.write(buf[cur].data(), buf[cur].size());
        }
    }
    // Only the first mk_bloom_window bytes of the table can be non-zero for this k; the rest of
    // the reference's 2^b / 8 bytes is written as zeros without ever being materialised.
    const uint64_t bloom_bytes = bloom_bits / 8;
    const uint64_t window = min<uint64_t>(bloom_bytes, mk_bloom_window(ix.shard[0]));
    vector<uint8_t> bloom(window);
    vector<uint64_t> gs(n);
    vector<uint32_t> ss(n);
    for (size_t r = 0; r < ix.shard.size(); ++r)
        if (mk_index_export(ix.shard[r], nullptr, gs.data() + ix.first[r], r == 0 ? bloom.data() : nullptr,
                            r == 0 ? bloom.size() : 0, ss.data() + ix.first[r]) != MK_OK)
            die(ix.shard[r], "mk_index_export");
    w.write(gs.data(), gs.size() * 8);
    w.write(bloom.data(), bloom.size());
    w.write_zeros(bloom_bytes - window);
    w.write(ss.data(), ss.size() * 4);
    w.close();                                  // throws on a short write: no silent truncation
    // file_names is not part of the dump (quirk G4: the reference's `-i ... -e` crashes).  A
    // side-car `<dump>.names` keeps them, one path per id, under a line that binds it to this
    // dump; the dump itself stays byte-compatible.  Without names a stale side-car is removed.
    const string side = path + ".names";
    if (ix.file_names.size() == n && n) {
        ofstream names(side.c_str());
        names << kNamesMagic << names_token(ix.k, ix.h, ix.b, ix.threshold, ss, gs) << "\n";
        for (const string& s : ix.file_names) names << s << "\n";
        names.close();
        if (!names) throw runtime_error("cannot write " + side);
    } else {
        remove(side.c_str());
    }
}

bool load_disk(Index& ix, const string& path) {
    if (!exists_test(path)) {
        cout << "File problem" << endl;                          // :683-686
        return false;
    }
    mkcli::LineReader in(path, ix.threads);     // our own dumps inflate `threads` members at a time
    unsigned char head[39];
    if (in.read(head, 39) != 39) {
        cerr << "miekki: truncated index dump" << endl;
        return false;
    }
    uint32_t n;
    uint64_t bloom_bits;
    memcpy(&ix.k, head + 0, 4);
    memcpy(&ix.h, head + 4, 4);
    memcpy(&ix.nbm, head + 8, 4);
    memcpy(&ix.nbmant, head + 12, 4);
    memcpy(&n, head + 16, 4);
    memcpy(&ix.b, head + 20, 4);
    memcpy(&bloom_bits, head + 24, 8);
    memcpy(&ix.threshold, head + 34, 4);
    if ((uint32_t)ix.gpus > max(1u, n)) ix.gpus = (int)max(1u, n);
    ix.create_shards();
    const uint64_t B = 1ull << ix.h;
    const size_t R = ix.shard.size();
    auto col = [&](size_t r) { return (uint32_t)(r * (uint64_t)n / R); };
    for (size_t r = 0; r < R; ++r)
        if (mk_index_import_begin(ix.shard[r], col(r + 1) - col(r)) != MK_OK) die(ix.shard[r], "mk_index_import_begin");
    bool ok = true;
    if (n) {
        const uint64_t slab = max<uint64_t>(1, min<uint64_t>(B, SLAB_BYTES / n));
        vector<uint8_t> buf[2];
        auto parse = [&](uint64_t r0, vector<uint8_t>& dst) {
            const uint64_t nr = min(slab, B - r0);
            dst.resize(nr * n);
            return in.read(dst.data(), dst.size()) == dst.size();
        };
        future<bool> next = async(launch::async, parse, (uint64_t)0, ref(buf[0]));
        int cur = 0;
        for (uint64_t r0 = 0; r0 < B && ok; r0 += slab, cur ^= 1) {
            ok = next.get();
            if (!ok) break;
            if (r0 + slab < B) next = async(launch::async, parse, r0 + slab, ref(buf[cur ^ 1]));
            const uint64_t nr = min(slab, B - r0);
            #pragma omp parallel for num_threads((int)R) schedule(static, 1)
            for (size_t r = 0; r < R; ++r)
                if (mk_index_import_rows(ix.shard[r], r0, nr, buf[cur].data() + col(r), n) != MK_OK)
                    die(ix.shard[r], "mk_index_import_rows");
        }
    }
    // of the table's 2^b / 8 bytes only the first mk_bloom_window can be non-zero for this k: the
    // rest is read past, not kept
    const uint64_t bloom_bytes = bloom_bits / 8;
    vector<uint8_t> bloom((size_t)min<uint64_t>(bloom_bytes, mk_bloom_window(ix.shard[0])));
    vector<uint64_t> gs(n);
    vector<uint32_t> ss(n);
    ok = ok && in.read(gs.data(), gs.size() * 8) == gs.size() * 8;
    ok = ok && in.read(bloom.data(), bloom.size()) == bloom.size();
    {
        vector<uint8_t> skip(16u << 20);
        for (uint64_t left = bloom_bytes - bloom.size(); ok && left > 0;) {
            const size_t m = (size_t)min<uint64_t>(left, skip.size());
            ok = in.read(skip.data(), m) == m;
            left -= m;
        }
    }
    ok = ok && in.read(ss.data(), ss.size() * 4) == ss.size() * 4;
    if (!ok) {
        cerr << "miekki: truncated index dump" << endl;
        return false;
    }
    for (size_t r = 0; r < R; ++r)
        if (mk_index_import_end(ix.shard[r], gs.data() + col(r), bloom.data(), bloom.size(), ss.data() + col(r)) != MK_OK)
            die(ix.shard[r], "mk_index_import_end");
    ix.seal_shards();      // every shard already holds the whole Bloom table: the fold is a no-op
    ix.compressed_flag = false;                                   // :705
    if (exists_test(path + ".names")) {      // side-car written by our -d: makes `-i ... -e` work
        mkcli::LineReader side(path + ".names");
        string first;
        side.getline(first);
        const string want = string(kNamesMagic) + names_token(ix.k, ix.h, ix.b, ix.threshold, ss, gs);
        vector<string> names;
        string name;
        while (!side.eof()) {
            side.getline(name);
            if (!name.empty()) names.push_back(name);
        }
        if (first == want && names.size() == n) ix.file_names = names;
        else cerr << "miekki: " << path << ".names does not belong to this dump: ignored" << endl;
    }
    return true;
}

// ---- merge: Miekki::merge_indexes, Miekki.cpp:901-910 ----------------------------------------
// `-i a.gz -i b.gz ...`: the genomes of each further dump follow those already loaded, ids in
// the order of the command line; the Bloom tables are folded (the reference's TODO at :907).
// The result is the index one in-order build of all the genome lists would have given.
bool merge_disk(Index& ix, const string& path) {
    Index other;
    other.threads = ix.threads;
    other.device0 = ix.device0;
    other.gpus = ix.gpus;
    other.devices = ix.devices;
    if (!load_disk(other, path)) return false;
    if (other.k != ix.k || other.h != ix.h || other.b != ix.b || other.nbm != ix.nbm) {
        cerr << "miekki: " << path << " was built with other -k / -h / -f / -b values: cannot merge" << endl;
        return false;
    }
    const uint32_t n0 = ix.size(), n1 = other.size();
    if (ix.shard.size() == 1 && other.shard.size() == 1) {
        if (mk_index_merge(ix.shard[0], other.shard[0]) != MK_OK) die(ix.shard[0], "mk_index_merge");
        mk_destroy(other.shard[0]);
    } else {
        // sharded over GPUs: the new columns stay where they were loaded, as further shards with
        // the following ids; seal_shards folds the Bloom tables in shard (= id) order
        ix.shard.insert(ix.shard.end(), other.shard.begin(), other.shard.end());
    }
    if (ix.file_names.size() == n0 && other.file_names.size() == n1)
        ix.file_names.insert(ix.file_names.end(), other.file_names.begin(), other.file_names.end());
    else
        ix.file_names.clear();
    ix.seal_shards();
    return true;
}

// ---- query: Miekki::query_file, Miekki.cpp:426-483 -------------------------------------------
struct ReadBatch {
    vector<string> heads, seqs;
    void clear() { heads.clear(); seqs.clear(); }
    size_t size() const { return seqs.size(); }
};

// the 2-line record loop of :458-475; fills up to `cap` reads, false when the file is done
bool next_reads(mkcli::LineReader& in, uint32_t k, size_t cap, ReadBatch& b, bool exact_filter) {
    b.clear();
    string head, ref;
    while (!in.eof() && b.size() < cap) {
        in.getline(head);
        in.getline(ref);
        if (ref.size() < k) continue;                            // :465
        if (exact_filter) {                                      // :736
            const char c = ref[0];
            if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != 'N') continue;
        }
        b.heads.push_back(head);
        b.seqs.push_back(ref);
    }
    return b.size() > 0;
}

void run_query(Index& ix, ReadBatch& b, uint32_t nresults, uint32_t min_score, double min_int,
               vector<mk_hit>& hits, vector<uint32_t>& nhits) {
    const size_t n = b.size();
    vector<const char*> ptr(n);
    vector<uint64_t> len(n);
    for (size_t i = 0; i < n; ++i) {
        ptr[i] = b.seqs[i].data();
        len[i] = b.seqs[i].size();
    }
    hits.assign(n * nresults, mk_hit{});
    nhits.assign(n, 0);
    const size_t R = ix.shard.size();
    if (R == 1) {
        if (mk_query(ix.shard[0], ptr.data(), len.data(), (uint32_t)n, nresults, min_score, min_int, hits.data(),
                     nhits.data()) != MK_OK)
            die(ix.shard[0], "mk_query");
        return;
    }
    // every shard sketches the reads and scans its own columns concurrently ...
    vector<mk_batch*> up(R, nullptr);
    #pragma omp parallel for num_threads((int)R) schedule(static, 1)
    for (size_t r = 0; r < R; ++r) {
        if (mk_batch_upload(ix.shard[r], ptr.data(), len.data(), (uint32_t)n, &up[r]) != MK_OK)
            die(ix.shard[r], "mk_batch_upload");
        if (mk_scan(ix.shard[r], up[r]) != MK_OK) die(ix.shard[r], "mk_scan");
    }
    // ... then the bounded heap walks the shards in ascending id order (Miekki.cpp:379-394)
    for (size_t r = 0; r < R; ++r) {
        if (mk_topk(ix.shard[r], nresults, min_score, min_int, hits.data(), nhits.data(), r > 0, r + 1 == R) != MK_OK)
            die(ix.shard[r], "mk_topk");
        mk_batch_free(ix.shard[r], up[r]);
    }
}

string hit_text(const mk_hit* h, uint32_t n) {                   // Miekki.cpp:442
    string s;
    for (uint32_t j = 0; j < n; ++j)
        s += to_string(h[j].genome) + "\t" + to_string(h[j].matches) + "\t" +
             to_string((unsigned)h[j].intersection) + "\t" + to_string(h[j].jaccard) + ";";
    return s;
}

void query_file(Index& ix, const string& path) {
    if (!exists_test(path)) {
        cout << "File problem" << endl;
        return;
    }
    mkcli::LineReader in(path);
    // Three stages overlap: while batch i is on the GPU, batch i+1 is parsed from the file and
    // the lines of batch i-1 are formatted and written (in order: one writer at a time).
    static const size_t cap = [] {
        const char* e = getenv("MIEKKI_QUERY_BATCH_READS");      // tests shrink it to force many batches
        return e ? max<size_t>(1, strtoull(e, nullptr, 10)) : (size_t)1 << 16;
    }();
    struct Scored {
        ReadBatch b;
        vector<mk_hit> hits;
        vector<uint32_t> nhits;
    };
    auto emit = [&ix](Scored* s) {
        vector<string> lines(s->b.size());
        #pragma omp parallel for num_threads(ix.threads) schedule(static)
        for (size_t i = 0; i < s->b.size(); ++i)
            lines[i] = s->b.heads[i] + ":" + hit_text(&s->hits[i * 10], s->nhits[i]) + "\n";   // :440-444
        for (const string& l : lines) *ix.out << l;
        delete s;
    };
    future<void> writer;
    Scored* cur = new Scored();
    bool have = next_reads(in, ix.k, cap, cur->b, false);
    while (have) {
        Scored* nxt = new Scored();
        future<bool> parsed = async(launch::async, [&in, &ix, nxt] { return next_reads(in, ix.k, cap, nxt->b, false); });
        cout << "-" << flush;                                    // :345
        run_query(ix, cur->b, 10, 10, 0.5 * ix.threshold, cur->hits, cur->nhits);   // :437
        if (writer.valid()) writer.get();
        writer = async(launch::async, emit, cur);
        have = parsed.get();
        cur = nxt;
    }
    delete cur;
    if (writer.valid()) writer.get();
    *ix.out << flush;
}

// ---- exact mode: Miekki::query_file_exact + ground_truth_batch, Miekki.cpp:723-759, 792-859 ---
struct Candidate {
    string seq, head;
    double jaccard, intersection;
};

// Miekki.cpp:823-859 for the genome files of one wave that went to one GPU: their records
// (:801-822) are already parsed; one mk_exact_many call, then the text block of each genome.
struct GenomeJob {
    const string* file;
    const vector<string>* recs;
    vector<Candidate>* cand;
    string* text;
};
void ground_truth_many(mk_ctx* ctx, vector<GenomeJob>& jobs) {
    if (jobs.empty()) return;
    vector<const char*> rp, qp;
    vector<uint64_t> rl, ql;
    vector<uint32_t> rc, qc;
    for (const GenomeJob& j : jobs) {
        for (const string& r : *j.recs) { rp.push_back(r.data()); rl.push_back(r.size()); }
        for (const Candidate& c : *j.cand) { qp.push_back(c.seq.data()); ql.push_back(c.seq.size()); }
        rc.push_back((uint32_t)j.recs->size());
        qc.push_back((uint32_t)j.cand->size());
    }
    vector<uint64_t> inter(qp.size()), uni(qp.size());
    if (mk_exact_many(ctx, (uint32_t)jobs.size(), rp.data(), rl.data(), rc.data(), qp.data(), ql.data(), qc.data(),
                      inter.data(), uni.data(), nullptr) != MK_OK)
        die(ctx, "mk_exact_many");
    size_t q = 0;
    for (const GenomeJob& j : jobs) {
        ostringstream os;
        for (const Candidate& c : *j.cand) {
            const double nb_inter = (double)inter[q], nb_union = (double)uni[q];
            ++q;
            if (nb_inter > 0) {                                  // :843
                const double real_jax = nb_inter / nb_union;     // :845-846
                os << real_jax << "\t" << c.jaccard << "\t" << nb_inter << "\t" << c.intersection << "\t" << c.head
                   << "\t" << *j.file << "\n";                   // :853
            }
        }
        *j.text = os.str();
    }
}

// Genomes are independent.  Their files are read and cut into records by the host team a wave
// ahead of the device (the reference re-reads the file inside ground_truth_batch, :800-822), and
// one host thread per GPU works through the parsed wave.
void ground_truth_all(Index& ix, map<uint32_t, vector<Candidate>>& per_genome) {
    vector<pair<uint32_t, vector<Candidate>*>> work;
    for (auto& kv : per_genome)
        work.push_back({kv.first, &kv.second});
    vector<string> text(work.size());
    const int R = (int)ix.shard.size();
    const size_t wave = max<size_t>(16, 4 * (size_t)ix.threads);
    struct Parsed {
        vector<vector<string>> recs;
        vector<char> ok;
    };
    auto parse = [&](size_t w0) {
        Parsed p;
        const size_t m = min(wave, work.size() - w0);
        p.recs.resize(m);
        p.ok.assign(m, 0);
        LoopError err;
        #pragma omp parallel for num_threads(ix.threads) schedule(dynamic, 1)
        for (size_t i = 0; i < m; ++i) {
            if (work[w0 + i].second->empty()) continue;
            const string& file = ix.file_names[work[w0 + i].first];
            if (!exists_test(file)) {
                #pragma omp critical(msg)
                cout << "File problem: " << file << endl;        // :796-799
                continue;
            }
            err.run([&] {
                p.recs[i] = mkcli::read_genome_records(file, ix.k);
                p.ok[i] = 1;
            });
        }
        err.rethrow();
        return p;
    };
    future<Parsed> next;
    if (!work.empty()) next = async(launch::async, parse, (size_t)0);
    for (size_t w0 = 0; w0 < work.size(); w0 += wave) {
        Parsed cur = next.get();
        if (w0 + wave < work.size()) next = async(launch::async, parse, w0 + wave);
        // the wave's genomes dealt round-robin to the GPUs, one call per GPU
        vector<vector<GenomeJob>> jobs((size_t)R);
        for (size_t i = 0, n_ok = 0; i < cur.recs.size(); ++i)
            if (cur.ok[i] && !work[w0 + i].second->empty())
                jobs[n_ok++ % (size_t)R].push_back({&ix.file_names[work[w0 + i].first], &cur.recs[i], work[w0 + i].second,
                                                    &text[w0 + i]});
        LoopError err;
        #pragma omp parallel for num_threads(R) schedule(static, 1)
        for (int r = 0; r < R; ++r) err.run([&] { ground_truth_many(ix.shard[(size_t)r], jobs[(size_t)r]); });
        err.rethrow();
    }
    for (const string& s : text) *ix.out << s;
    per_genome.clear();
}

void query_file_exact(Index& ix, const string& path) {
    if (!exists_test(path)) {
        cout << "File problem" << endl;
        return;
    }
    if (ix.file_names.empty()) {
        // the reference segfaults here (file_names is not in the dump, quirk G4)
        cerr << "miekki: exact mode needs the genome list (-l) in the same run" << endl;
        return;
    }
    mkcli::LineReader in(path);
    ReadBatch b, nxt;
    vector<mk_hit> hits;
    vector<uint32_t> nhits;
    map<uint32_t, vector<Candidate>> per_genome;
    size_t held = 0;
    bool have = next_reads(in, ix.k, 1 << 14, b, true);
    while (have) {
        // the next batch is parsed while this one is scored (and its genomes intersected)
        future<bool> parsed = async(launch::async, [&in, &ix, &nxt] { return next_reads(in, ix.k, 1 << 14, nxt, true); });
        run_query(ix, b, 5, 10, (double)ix.threshold, hits, nhits);   // :741
        for (size_t i = 0; i < b.size(); ++i)
            for (uint32_t j = 0; j < nhits[i]; ++j) {
                const mk_hit& sim = hits[i * 5 + j];
                per_genome[sim.genome].push_back({b.seqs[i], b.heads[i], sim.jaccard, sim.intersection});
                ++held;
            }
        if (held > (1u << 20)) {                 // the reference flushes per genome at 100 (:745)
            ground_truth_all(ix, per_genome);
            held = 0;
        }
        have = parsed.get();
        swap(b.heads, nxt.heads);
        swap(b.seqs, nxt.seqs);
    }
    ground_truth_all(ix, per_genome);
    *ix.out << flush;
}

// ---- -A: whole-file queries, Miekki.cpp:487-514 + :592-612 ------------------------------------
// Each listed file is one query: all non-header lines concatenated; a line `path:hits` is
// written only when there are hits.  Files are parsed in parallel and scored in waves; lines
// come out in list order (the reference's order for -t 1).
void query_file_of_file(Index& ix, const string& list) {
    if (!exists_test(list)) {
        cout << "Missed file of file: " << list << endl;
        return;
    }
    const vector<string> names = read_list(list);
    const size_t wave = max<size_t>(16, 2 * (size_t)ix.threads);
    struct Wave {
        vector<string> seqs;
        vector<char> ok;
    };
    auto parse = [&](size_t w0) {
        Wave w;
        const size_t m = min(wave, names.size() - w0);
        w.seqs.resize(m);
        w.ok.assign(m, 0);
        LoopError err;
        #pragma omp parallel for num_threads(ix.threads) schedule(dynamic, 1)
        for (size_t i = 0; i < m; ++i) {
            const string& fn = names[w0 + i];
            if (!exists_test(fn)) {
                #pragma omp critical(msg)
                cout << "File problem" << endl;                  // :488
                continue;
            }
            err.run([&] {
                w.seqs[i] = mkcli::read_genome_concat(fn);       // :491-497
                w.ok[i] = w.seqs[i].size() >= ix.k;              // :498
            });
        }
        err.rethrow();
        return w;
    };
    // wave i+1 is read and parsed while wave i is scored
    future<Wave> next;
    if (!names.empty()) next = async(launch::async, parse, (size_t)0);
    for (size_t w0 = 0; w0 < names.size(); w0 += wave) {
        Wave cur = next.get();
        if (w0 + wave < names.size()) next = async(launch::async, parse, w0 + wave);
        const size_t m = cur.seqs.size();
        ReadBatch b;
        for (size_t i = 0; i < m; ++i) {
            if (cur.ok[i]) {
                b.heads.push_back(names[w0 + i]);
                b.seqs.push_back(std::move(cur.seqs[i]));
            }
            cout << "-" << flush;                                // :608
        }
        if (b.size() == 0) continue;
        vector<mk_hit> hits;
        vector<uint32_t> nhits;
        run_query(ix, b, 10, 10, 0.5 * ix.threshold, hits, nhits);   // :500
        for (size_t i = 0; i < b.size(); ++i)
            if (nhits[i])                                        // :506
                *ix.out << b.heads[i] << ":" << hit_text(&hits[i * 10], nhits[i]) << "\n";   // :509
    }
    *ix.out << flush;
}

// ---- -A -e: Miekki.cpp:616-645 + :763-788 -------------------------------------------------
// The query is the file's lines after the first that are at least k long and start with one of
// ACGTN (shorter lines are dropped, :771); candidates = filter_results(..., 5, 5, threshold).
void query_file_of_file_exact(Index& ix, const string& list) {
    if (!exists_test(list)) {
        cout << "Missed file of file: " << list << endl;
        return;
    }
    if (ix.file_names.empty()) {
        cerr << "miekki: exact mode needs the genome list (-l) in the same run" << endl;
        return;
    }
    const vector<string> names = read_list(list);
    map<uint32_t, vector<Candidate>> per_genome;
    ReadBatch b;
    for (const string& fn : names) {
        cout << "-" << flush;
        if (!exists_test(fn)) {
            cout << "File problem" << endl;
            continue;
        }
        mkcli::LineReader in(fn);
        string ref, head, line;
        in.getline(head);                                        // :768
        while (!in.eof()) {
            in.getline(line);
            if (line.size() < ix.k) continue;                    // :771
            const char c = line[0];
            if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != 'N') continue;
            ref += line;
        }
        if (ref.size() <= ix.k) continue;       // the reference sketches nothing here: no hits
        b.heads.push_back(head);
        b.seqs.push_back(std::move(ref));
    }
    if (b.size()) {
        vector<mk_hit> hits;
        vector<uint32_t> nhits;
        run_query(ix, b, 5, 5, (double)ix.threshold, hits, nhits);   // :778
        for (size_t i = 0; i < b.size(); ++i)
            for (uint32_t j = 0; j < nhits[i]; ++j) {
                const mk_hit& sim = hits[i * 5 + j];
                per_genome[sim.genome].push_back({b.seqs[i], b.heads[i], sim.jaccard, sim.intersection});
            }
    }
    ground_truth_all(ix, per_genome);
    *ix.out << flush;
}

}  // namespace

static int run(int argc, char** argv);

// A corrupt or truncated input, an unwritable dump: one line on stderr and exit code 1, never
// std::terminate and never "The end" after a failed step.
int main(int argc, char** argv) {
    try {
        return run(argc, argv);
    } catch (const exception& e) {
        cerr << "miekki: " << e.what() << endl;
        return 1;
    }
}

static int run(int argc, char** argv) {
    if (argc < 2) {
        help();
        exit(0);
    }
    string index_file, list_file, query_fa, query_list, output_file("out.txt"), index_dump;
    vector<string> index_files;                  // every -i, in order (the first is loaded, the rest merged)
    uint64_t H = 17, core_number = 8, kmer_size = 31, bloom_size = 33, fingerprint_size = 3;
    double threshold = 200;
    bool exact_mode = false;
    int device = 0, gpus = 1;
    vector<int> devices;
    static const option longopts[] = {{"device", required_argument, nullptr, 1000},
                                      {"gpus", required_argument, nullptr, 1001},
                                      {"devices", required_argument, nullptr, 1002},
                                      {nullptr, 0, nullptr, 0}};
    int c;
    // string-valued and numeric flags are bound to their variables once; the loop is generic
    const map<int, string*> text_flags = {{'i', &index_file}, {'l', &list_file},  {'a', &query_fa},
                                          {'A', &query_list}, {'o', &output_file}, {'d', &index_dump}};
    const map<int, uint64_t*> count_flags = {{'h', &H}, {'t', &core_number}, {'k', &kmer_size},
                                             {'f', &fingerprint_size}, {'b', &bloom_size}};
    while ((c = getopt_long(argc, argv, "i:l:a:h:t:f:k:s:b:o:ed:A:", longopts, nullptr)) != -1) {
        if (auto t = text_flags.find(c); t != text_flags.end()) {
            *t->second = optarg;
            if (c == 'i') index_files.push_back(optarg);
        } else if (auto n = count_flags.find(c); n != count_flags.end()) {
            *n->second = (uint64_t)stoi(optarg);
        } else if (c == 's') {
            threshold = stof(optarg);
        } else if (c == 'e') {
            exact_mode = true;
        } else if (c == 1000) {
            device = stoi(optarg);
        } else if (c == 1001) {
            gpus = max(1, stoi(optarg));
        } else if (c == 1002) {
            stringstream ss(optarg);
            string tok;
            while (getline(ss, tok, ',')) devices.push_back(stoi(tok));
        }
    }
    const uint32_t bit_per_min = (uint32_t)(5 + fingerprint_size);
    cout << "Using " << bit_per_min << " bits per minimizer, " << int_to_string(1ull << H) << " minimizers so "
         << int_to_string((uint64_t)bit_per_min * (1ull << H)) << " bits per sequences" << endl;   // main.cpp:186
    auto start = chrono::system_clock::now();
    Index ix;
    ix.threads = (int)max<uint64_t>(1, core_number);
    ix.device0 = device;
    ix.gpus = devices.empty() ? gpus : (int)devices.size();
    ix.devices = devices;
    if (!index_file.empty()) {
        if (!load_disk(ix, index_files[0])) return 1;
        for (size_t i = 1; i < index_files.size(); ++i)
            if (!merge_disk(ix, index_files[i])) return 1;
        ix.out = new ofstream(output_file.c_str());
        cout << "I output results in " << output_file << endl;
        cout << "Load sucessful" << endl;                        // main.cpp:193 (sic)
    } else if (!list_file.empty()) {
        ix.k = (uint32_t)kmer_size;
        ix.h = (uint32_t)H;
        ix.nbm = bit_per_min;
        ix.b = (uint32_t)bloom_size;
        ix.threshold = (uint32_t)threshold;                      // double -> uint32_t, Miekki.h:66
        if (bit_per_min != 8) {
            cout << "not implemented" << endl;                   // Miekki.cpp:236-237 (quirk G12)
            exit(0);
        }
        {
            PhaseClock pc("create_shards");
            ix.create_shards();
        }
        ix.out = new ofstream(output_file.c_str());
        cout << "I output results in " << output_file << endl;   // Miekki.h:75
        index_file_of_file(ix, list_file);
        ix.compressed_flag = true;      // the reference has run compress_index(1) here (main.cpp:198)
    } else {
        cout << "What am I supposed to index ? use either -i or -l options please" << endl;
        help();
        exit(0);
    }
    if (!index_dump.empty()) {
        cout << "I write this index on the disk for later" << endl;
        dump_disk(ix, index_dump);
    }
    auto end_index = chrono::system_clock::now();
    chrono::duration<double> elapsed = end_index - start;
    cout << "elapsed time: " << elapsed.count() << "s\n";
    if (!query_fa.empty()) {
        if (exact_mode) {
            cout << "running in exact mode, actual intersection will be computed on hits found by the index" << endl;
            query_file_exact(ix, query_fa);
        } else {
            cout << "running in approx mode, intersection is estimated by the index" << endl;
            query_file(ix, query_fa);
        }
    } else if (!query_list.empty()) {
        if (exact_mode) {
            cout << "running in exact mode, actual intersection will be computed on hits found by the index" << endl;
            query_file_of_file_exact(ix, query_list);
        } else {
            cout << "running in approx mode, intersection is estimated by the index" << endl;
            query_file_of_file(ix, query_list);
        }
    } else {
        cout << "No query file, No queries" << endl;
    }
    auto end_query = chrono::system_clock::now();
    elapsed = end_query - end_index;
    cout << "elapsed time: " << elapsed.count() << "s\n";
    cout << "The end" << endl;
    if (ix.out) ix.out->close();
    for (mk_ctx* s : ix.shard) mk_destroy(s);
    return 0;
}
