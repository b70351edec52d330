// fasta.hpp -- host-side text ingestion for the miekki CLI: a line reader over plain,
// gzip or zlib files (what zstr::ifstream gives the reference, zstr.hpp:136-209) and the
// reference's three ways of cutting a file into sequences.
#pragma once
#include <zlib.h>

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
    explicit LineReader(const std::string& path) : in_(1 << 20), out_(1 << 20) {
        f_ = fopen(path.c_str(), "rb");
        if (!f_) throw std::runtime_error("cannot open " + path);
        memset(&zs_, 0, sizeof(zs_));
        fill_in();
        // zstr.hpp:154-167: gzip (1f 8b) or zlib (78 01/9c/da) header, else plain text
        const unsigned char* p = in_.data();
        compressed_ = in_len_ >= 2 && ((p[0] == 0x1F && p[1] == 0x8B) ||
                                       (p[0] == 0x78 && (p[1] == 0x01 || p[1] == 0x9C || p[1] == 0xDA)));
        if (compressed_) {
            if (inflateInit2(&zs_, 15 + 32) != Z_OK) throw std::runtime_error("inflateInit2 failed");
            zs_.next_in = in_.data();
            zs_.avail_in = (uInt)in_len_;
        }
    }
    ~LineReader() {
        if (compressed_) inflateEnd(&zs_);
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
    const unsigned char* cur() const { return compressed_ ? out_.data() : in_.data(); }
    void fill_in() {
        in_len_ = fread(in_.data(), 1, in_.size(), f_);
    }
    bool refill() {
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
};

// gzip level-1 writer (zstr::ofstream, zstr.hpp:82,230)
class GzWriter {
public:
    explicit GzWriter(const std::string& path) : buf_(1 << 20) {
        f_ = fopen(path.c_str(), "wb");
        if (!f_) throw std::runtime_error("cannot open " + path);
        memset(&zs_, 0, sizeof(zs_));
        if (deflateInit2(&zs_, 1, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK)
            throw std::runtime_error("deflateInit2 failed");
    }
    ~GzWriter() { close(); }
    void write(const void* p, size_t n) {
        const unsigned char* src = (const unsigned char*)p;
        while (n) {
            const size_t m = std::min<size_t>(n, 1u << 30);
            zs_.next_in = const_cast<unsigned char*>(src);
            zs_.avail_in = (uInt)m;
            pump(Z_NO_FLUSH);
            src += m;
            n -= m;
        }
    }
    void close() {
        if (!f_) return;
        zs_.next_in = nullptr;
        zs_.avail_in = 0;
        pump(Z_FINISH);
        deflateEnd(&zs_);
        fclose(f_);
        f_ = nullptr;
    }

private:
    void pump(int flush) {
        for (;;) {
            zs_.next_out = buf_.data();
            zs_.avail_out = (uInt)buf_.size();
            const int r = deflate(&zs_, flush);
            fwrite(buf_.data(), 1, buf_.size() - zs_.avail_out, f_);
            if (flush == Z_FINISH) {
                if (r == Z_STREAM_END) return;
            } else if (zs_.avail_in == 0 && zs_.avail_out != 0) {
                return;
            }
        }
    }
    FILE* f_ = nullptr;
    z_stream zs_;
    std::vector<unsigned char> buf_;
};

// Miekki.cpp:559-567 (index_file_of_file): every line that does not start with '>' is
// appended; records are concatenated without separator (quirk G9).
inline std::string read_genome_concat(const std::string& path) {
    LineReader in(path);
    std::string ref, line;
    while (!in.eof()) {
        in.getline(line);
        if (line.empty() || line[0] != '>') ref += line;
    }
    return ref;
}

// Miekki.cpp:801-822 (ground_truth_batch): the strings whose k-mers form set B.  At a header
// line the record is flushed only if it has at least k characters; a shorter one is kept and
// runs into the next record (quirk G16).
inline std::vector<std::string> read_genome_records(const std::string& path, uint32_t k) {
    LineReader in(path);
    std::vector<std::string> recs;
    std::string ref, line;
    while (!in.eof()) {
        in.getline(line);
        if (!line.empty() && line[0] == '>') {
            if (ref.size() >= k) {
                recs.push_back(std::move(ref));
                ref.clear();
            }
        } else {
            ref += line;
        }
    }
    if (ref.size() >= k) recs.push_back(std::move(ref));
    return recs;
}

}  // namespace mkcli
