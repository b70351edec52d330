// fasta.hpp -- host-side text ingestion for the miekki CLI: a line reader over plain,
// gzip or zlib files (what zstr::ifstream gives the reference, zstr.hpp:136-209) and the
// reference's three ways of cutting a file into sequences.
#pragma once
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace mkcli {

// Sequential reader with std::getline semantics on the decompressed bytes.
class LineReader {
public:
    explicit LineReader(const std::string& path, int threads = 1) : in_(1 << 20), out_(1 << 20) {
        f_ = fopen(path.c_str(), "rb");
        if (!f_) throw std::runtime_error("cannot open " + path);
        memset(&zs_, 0, sizeof(zs_));
        fill_in();
        // zstr.hpp:154-167: gzip (1f 8b) or zlib (78 01/9c/da) header, else plain text
        const unsigned char* p = in_.data();
        compressed_ = in_len_ >= 2 && ((p[0] == 0x1F && p[1] == 0x8B) ||
                                       (p[0] == 0x78 && (p[1] == 0x01 || p[1] == 0x9C || p[1] == 0xDA)));
        // our own dumps: gzip members that carry their sizes (ParallelGzWriter) are inflated
        // `threads` at a time
        indexed_ = threads > 1 && in_len_ >= 24 && member_sizes(p, nullptr, nullptr);
        if (indexed_) {
            threads_ = threads;
            fseek(f_, 0, SEEK_SET);
            return;
        }
        if (compressed_) {
            if (inflateInit2(&zs_, 15 + 32) != Z_OK) throw std::runtime_error("inflateInit2 failed");
            zs_.next_in = in_.data();
            zs_.avail_in = (uInt)in_len_;
        }
    }
    ~LineReader() {
        if (compressed_ && !indexed_) inflateEnd(&zs_);
        if (f_) fclose(f_);
    }
    LineReader(const LineReader&) = delete;
    LineReader& operator=(const LineReader&) = delete;

    // true once a read hit the end of the data (like istream::eof after a getline)
    bool eof() const { return eof_; }

    // std::getline: extracts up to '\n' (dropped).  Returns false only if nothing could be
    // extracted at all (stream already exhausted); the line is cleared either way.
    bool getline(std::string& line) {
        line.clear();
        bool got_any = false;
        for (;;) {
            if (pos_ == len_) {
                if (!refill()) {
                    eof_ = true;
                    return got_any;
                }
            }
            const unsigned char* base = cur() + pos_;
            const size_t avail = len_ - pos_;
            const void* nl = memchr(base, '\n', avail);
            if (nl) {
                const size_t n = (const unsigned char*)nl - base;
                line.append((const char*)base, n);
                pos_ += n + 1;
                return true;
            }
            line.append((const char*)base, avail);
            got_any = got_any || avail > 0;
            pos_ = len_;
        }
    }

    // The loop of Miekki.cpp:559-567 in one pass over the reader's buffer: every line that does
    // not start with '>' is appended to `ref` (one copy per byte instead of getline's two).
    void concat_sequence_lines(std::string& ref) {
        bool at_line_start = true, skipping = false;
        for (;;) {
            if (pos_ == len_ && !refill()) break;
            const unsigned char* base = cur() + pos_;
            const size_t avail = len_ - pos_;
            if (at_line_start) {
                skipping = base[0] == '>';
                at_line_start = false;
            }
            const void* nl = memchr(base, '\n', avail);
            const size_t seg = nl ? (size_t)((const unsigned char*)nl - base) : avail;
            if (!skipping) ref.append((const char*)base, seg);
            pos_ += seg + (nl ? 1 : 0);
            if (nl) at_line_start = true;
        }
        eof_ = true;
    }
    // The loop of Miekki.cpp:801-822 in one pass over the reader's buffer: lines are appended to the
    // current record; at a line that starts with '>' the record is closed if it holds at least k
    // characters (a shorter one is kept and runs into the next record, quirk G16).
    void split_records(std::vector<std::string>& recs, uint32_t k, size_t reserve_bytes) {
        std::string ref;
        ref.reserve(reserve_bytes);
        bool at_line_start = true, header = false;
        for (;;) {
            if (pos_ == len_ && !refill()) break;
            const unsigned char* base = cur() + pos_;
            const size_t avail = len_ - pos_;
            if (at_line_start) {
                header = base[0] == '>';
                at_line_start = false;
                if (header && ref.size() >= k) {
                    recs.push_back(std::move(ref));
                    ref.clear();
                }
            }
            const void* nl = memchr(base, '\n', avail);
            const size_t seg = nl ? (size_t)((const unsigned char*)nl - base) : avail;
            if (!header) ref.append((const char*)base, seg);
            pos_ += seg + (nl ? 1 : 0);
            if (nl) at_line_start = true;
        }
        eof_ = true;
        if (ref.size() >= k) recs.push_back(std::move(ref));
    }
    // bytes of the file on disk (a lower bound of the text for compressed files)
    size_t file_bytes() const {
        const long at = ftell(f_);
        if (at < 0 || fseek(f_, 0, SEEK_END) != 0) return 0;
        const long end = ftell(f_);
        fseek(f_, at, SEEK_SET);
        return end > 0 ? (size_t)end : 0;
    }
    bool is_compressed() const { return compressed_; }

    // raw bytes (dump loading)
    size_t read(void* dst, size_t n) {
        size_t done = 0;
        while (done < n) {
            if (pos_ == len_ && !refill()) {
                eof_ = true;
                break;
            }
            const size_t m = std::min(n - done, len_ - pos_);
            memcpy((char*)dst + done, cur() + pos_, m);
            pos_ += m;
            done += m;
        }
        return done;
    }

private:
    const unsigned char* cur() const {
        return indexed_ ? wave_[wave_pos_].data() : compressed_ ? out_.data() : in_.data();
    }
    // header of a member written by ParallelGzWriter -> its total and payload sizes
    static bool member_sizes(const unsigned char* h, uint32_t* total, uint32_t* usize) {
        if (!(h[0] == 0x1F && h[1] == 0x8B && h[2] == 8 && h[3] == 4 && h[10] == 12 && h[11] == 0 && h[12] == 'M' &&
              h[13] == 'K' && h[14] == 8 && h[15] == 0))
            return false;
        if (total) memcpy(total, h + 16, 4);
        if (usize) memcpy(usize, h + 20, 4);
        return true;
    }
    // indexed mode: read the next `threads_` members from the file, inflate them in parallel
    bool next_wave() {
        std::vector<std::vector<unsigned char>> raw;
        std::vector<uint32_t> usz;
        for (int i = 0; i < threads_; ++i) {
            unsigned char h[24];
            const size_t got = fread(h, 1, 24, f_);
            if (got == 0) break;
            uint32_t total = 0, usize = 0;
            if (got != 24 || !member_sizes(h, &total, &usize) || total < 32)
                throw std::runtime_error("index dump: gzip member without the size record");
            raw.emplace_back(total);
            memcpy(raw.back().data(), h, 24);
            if (fread(raw.back().data() + 24, 1, total - 24, f_) != total - 24)
                throw std::runtime_error("index dump: truncated gzip member");
            usz.push_back(usize);
        }
        wave_.assign(raw.size(), {});
        wave_pos_ = 0;
        if (raw.empty()) return false;
        int bad = 0;
        #pragma omp parallel for num_threads(threads_) schedule(dynamic, 1)
        for (size_t i = 0; i < raw.size(); ++i) {
            wave_[i].resize(usz[i]);
            z_stream zs;
            memset(&zs, 0, sizeof(zs));
            bool ok = inflateInit2(&zs, -15) == Z_OK;
            if (ok) {
                zs.next_in = raw[i].data() + 24;
                zs.avail_in = (uInt)(raw[i].size() - 24 - 8);
                zs.next_out = wave_[i].data();
                zs.avail_out = (uInt)usz[i];
                ok = inflate(&zs, Z_FINISH) == Z_STREAM_END && zs.avail_out == 0;
                inflateEnd(&zs);
            }
            uint32_t crc = 0;
            memcpy(&crc, raw[i].data() + raw[i].size() - 8, 4);
            ok = ok && crc == (uint32_t)crc32(crc32(0L, Z_NULL, 0), wave_[i].data(), (uInt)usz[i]);
            if (!ok) {
                #pragma omp atomic write
                bad = 1;
            }
        }
        if (bad) throw std::runtime_error("index dump: corrupt gzip member");
        return true;
    }
    void fill_in() {
        in_len_ = fread(in_.data(), 1, in_.size(), f_);
    }
    bool refill() {
        if (indexed_) {
            // members may be empty (an empty payload is still one member): skip them
            for (;;) {
                if (wave_started_ && wave_pos_ + 1 < wave_.size()) ++wave_pos_;
                else if (!next_wave()) return false;
                wave_started_ = true;
                pos_ = 0;
                len_ = wave_[wave_pos_].size();
                if (len_ > 0) return true;
            }
        }
        if (!compressed_) {
            if (first_plain_) {          // the constructor already read the first block
                first_plain_ = false;
            } else {
                fill_in();
            }
            pos_ = 0;
            len_ = in_len_;
            in_len_ = 0;
            return len_ > 0;
        }
        for (;;) {
            if (z_done_) return false;
            if (zs_.avail_in == 0) {
                fill_in();
                if (in_len_ == 0) return false;
                zs_.next_in = in_.data();
                zs_.avail_in = (uInt)in_len_;
            }
            zs_.next_out = out_.data();
            zs_.avail_out = (uInt)out_.size();
            const int r = inflate(&zs_, Z_NO_FLUSH);
            if (r == Z_STREAM_END) {
                // concatenated members: keep going if more input follows
                if (zs_.avail_in == 0) {
                    fill_in();
                    zs_.next_in = in_.data();
                    zs_.avail_in = (uInt)in_len_;
                }
                if (zs_.avail_in == 0) z_done_ = true;
                else inflateReset(&zs_);
            } else if (r != Z_OK && r != Z_BUF_ERROR) {
                throw std::runtime_error(std::string("inflate failed: ") + (zs_.msg ? zs_.msg : "?"));
            }
            pos_ = 0;
            len_ = out_.size() - zs_.avail_out;
            if (len_ > 0) return true;
            if (r == Z_BUF_ERROR && zs_.avail_in == 0 && feof(f_)) return false;
        }
    }

    FILE* f_ = nullptr;
    z_stream zs_;
    std::vector<unsigned char> in_, out_;
    size_t in_len_ = 0, pos_ = 0, len_ = 0;
    bool compressed_ = false, eof_ = false, z_done_ = false, first_plain_ = true;
    bool indexed_ = false, wave_started_ = false;
    int threads_ = 1;
    std::vector<std::vector<unsigned char>> wave_;
    size_t wave_pos_ = 0;
};

// Parallel gzip-1 writer for the index dump (zstr::ofstream writes gzip level 1,
// zstr.hpp:82,230): the payload is cut into 32 MiB chunks, each deflated by its own thread
// into an independent gzip member, written in order.  Concatenated
// members are a valid gzip file, and the reference's reader restarts its inflator at every
// member end (zstr.hpp:193-197), so the dump stays loadable by the reference's -i.
class ParallelGzWriter {
public:
    static constexpr size_t member_header_bytes() { return 24; }
    ParallelGzWriter(const std::string& path, int threads) : threads_(threads < 1 ? 1 : threads) {
        f_ = fopen(path.c_str(), "wb");
        if (!f_) throw std::runtime_error("cannot open " + path);
    }
    ~ParallelGzWriter() {
        try { close(); } catch (...) {}
    }
    // small pieces (header fields) are gathered and leave with the next large block
    void write(const void* p, size_t n) {
        const unsigned char* src = (const unsigned char*)p;
        if (n < CHUNK / 4) {
            pending_.insert(pending_.end(), src, src + n);
            if (pending_.size() >= CHUNK) flush_pending();
            return;
        }
        flush_pending();
        compress_and_write(src, n);
    }
    // n zero bytes (the never-touched tail of the Bloom table: 15/16 of the reference's 1 GiB at
    // the default -k / -b).  Whole chunks reuse one member deflated once.
    void write_zeros(size_t n) {
        flush_pending();
        if (n >= CHUNK && zero_member_.empty()) {
            const std::vector<unsigned char> z(CHUNK, 0);
            deflate_member(z.data(), CHUNK, zero_member_);
        }
        for (; n >= CHUNK; n -= CHUNK) {
            put(zero_member_.data(), zero_member_.size());
            wrote_any_ = true;
        }
        if (n) {
            const std::vector<unsigned char> z(n, 0);
            compress_and_write(z.data(), n);
        }
    }
    // Throws when any byte could not be written (full disk, I/O error): a dump must never be
    // silently short.
    void close() {
        if (!f_) return;
        FILE* f = f_;
        try {
            flush_pending();
            if (!wrote_any_) {                  // an empty payload is still a valid gzip file
                std::vector<unsigned char> out;
                deflate_member(nullptr, 0, out);
                put(out.data(), out.size());
            }
        } catch (...) {
            f_ = nullptr;
            fclose(f);
            throw;
        }
        f_ = nullptr;
        if (fclose(f) != 0) throw std::runtime_error("index dump: close failed (disk full?)");
    }

private:
    static constexpr size_t CHUNK = 32u << 20;
    void put(const void* p, size_t n) {
        if (n && fwrite(p, 1, n, f_) != n) throw std::runtime_error("index dump: short write (disk full?)");
    }
    // One gzip member written by hand around a raw deflate stream, so that its header can say
    // how long the member is (the way BGZF does): FEXTRA subfield 'M','K' = {u32 member bytes,
    // u32 payload bytes}.  Any gzip reader skips the field; IndexedGzSource hops from member to
    // member with it and inflates them in parallel.
    static void deflate_member(const unsigned char* src, size_t n, std::vector<unsigned char>& out) {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (deflateInit2(&zs, 1, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK)
            throw std::runtime_error("deflateInit2 failed");
        const size_t head = member_header_bytes();
        out.resize(head + deflateBound(&zs, (uLong)n) + 64 + 8);
        zs.next_in = const_cast<unsigned char*>(src);
        zs.avail_in = (uInt)n;
        zs.next_out = out.data() + head;
        zs.avail_out = (uInt)(out.size() - head - 8);
        const int r = deflate(&zs, Z_FINISH);
        const size_t produced = (out.size() - head - 8) - zs.avail_out;
        deflateEnd(&zs);
        if (r != Z_STREAM_END) throw std::runtime_error("deflate failed");
        const uint32_t total = (uint32_t)(head + produced + 8), usize = (uint32_t)n;
        const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), src, (uInt)n);
        unsigned char* h = out.data();
        const unsigned char fixed[12] = {0x1F, 0x8B, 8, 4 /* FEXTRA */, 0, 0, 0, 0, 0, 3 /* unix */, 12, 0 /* XLEN */};
        memcpy(h, fixed, 12);
        h[12] = 'M'; h[13] = 'K'; h[14] = 8; h[15] = 0;
        memcpy(h + 16, &total, 4);
        memcpy(h + 20, &usize, 4);
        memcpy(h + head + produced, &crc, 4);
        memcpy(h + head + produced + 4, &usize, 4);
        out.resize(total);
    }
    void flush_pending() {
        if (pending_.empty()) return;
        std::vector<unsigned char> tmp;
        tmp.swap(pending_);
        compress_and_write(tmp.data(), tmp.size());
    }
    void compress_and_write(const unsigned char* src, size_t n) {
        const size_t nchunks = (n + CHUNK - 1) / CHUNK;
        // waves of `threads_` chunks keep memory bounded
        for (size_t c0 = 0; c0 < nchunks; c0 += (size_t)threads_) {
            const size_t m = std::min<size_t>((size_t)threads_, nchunks - c0);
            std::vector<std::vector<unsigned char>> outs(m);
            int bad = 0;                        // an exception must not leave the parallel region
            #pragma omp parallel for num_threads(threads_) schedule(dynamic, 1)
            for (size_t i = 0; i < m; ++i) {
                const size_t off = (c0 + i) * CHUNK;
                try {
                    deflate_member(src + off, std::min(CHUNK, n - off), outs[i]);
                } catch (...) {
                    #pragma omp atomic write
                    bad = 1;
                }
            }
            if (bad) throw std::runtime_error("index dump: deflate failed");
            for (size_t i = 0; i < m; ++i) put(outs[i].data(), outs[i].size());
        }
        wrote_any_ = wrote_any_ || n > 0;
    }
    FILE* f_ = nullptr;
    int threads_;
    bool wrote_any_ = false;
    std::vector<unsigned char> pending_, zero_member_;
};

// Miekki.cpp:559-567 (index_file_of_file): every line that does not start with '>' is
// appended; records are concatenated without separator (quirk G9).
inline std::string read_genome_concat(const std::string& path) {
    LineReader in(path);
    std::string ref;
    const size_t bytes = in.file_bytes();
    ref.reserve(in.is_compressed() ? bytes * 4 : bytes);       // gzip-1 of DNA text: about 3.5x
    in.concat_sequence_lines(ref);
    return ref;
}

// Miekki.cpp:801-822 (ground_truth_batch): the strings whose k-mers form set B.  At a header
// line the record is flushed only if it has at least k characters; a shorter one is kept and
// runs into the next record (quirk G16).
inline std::vector<std::string> read_genome_records(const std::string& path, uint32_t k) {
    LineReader in(path);
    std::vector<std::string> recs;
    const size_t bytes = in.file_bytes();
    in.split_records(recs, k, in.is_compressed() ? bytes * 4 : bytes);
    return recs;
}

}  // namespace mkcli
