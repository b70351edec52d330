// Host-side text ingestion of the CLI (miekki_b200/cli/fasta.hpp), exercised without a GPU.
// usage: fasta_host_test <cmd> <args...>; prints results on stdout for tests/test_cli_host_cpu.py
#include <cstdint>
#include <iostream>
#include <string>

#include "../../miekki_b200/cli/fasta.hpp"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const std::string cmd = argv[1];
    if (cmd == "lines") {                       // getline loop exactly like the reference's while(!eof)
        mkcli::LineReader in(argv[2]);
        std::string line;
        size_t n = 0;
        while (!in.eof()) {
            in.getline(line);
            std::cout << n++ << ":" << line << "\n";
        }
    } else if (cmd == "concat") {               // Miekki.cpp:559-567
        std::cout << mkcli::read_genome_concat(argv[2]);
    } else if (cmd == "records") {              // Miekki.cpp:801-822, k = argv[3]
        for (const std::string& r : mkcli::read_genome_records(argv[2], (uint32_t)std::stoul(argv[3])))
            std::cout << r << "\n";
    } else if (cmd == "gzwrite" || cmd == "gzwrite-readonly") {   // ParallelGzWriter: argv[2] = file, argv[3] = payload bytes
        const size_t n = std::stoull(argv[3]);
        std::string payload(n, '\0');
        uint64_t x = 88172645463325252ull;
        for (size_t i = 0; i < n; ++i) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            payload[i] = (i / 4096) % 3 == 0 ? 0 : (char)(x & 0xFF);     // zero runs and noise
        }
        if (cmd == "gzwrite") {
            mkcli::ParallelGzWriter w(argv[2], 4);
            w.write("HEAD", 4);
            w.write(payload.data(), n / 3);
            w.write(payload.data() + n / 3, n - n / 3);
            w.write("TAIL", 4);
            w.close();
        }
        // read it back with our own reader: sequential inflate, or (argv[4] threads) member-parallel
        mkcli::LineReader in(argv[2], argc > 4 ? std::stoi(argv[4]) : 1);
        std::string back(n + 8, '\0');
        const size_t got = in.read(&back[0], n + 8);
        char extra;
        const bool at_end = in.read(&extra, 1) == 0;
        std::cout << (got == n + 8 && at_end && back.substr(0, 4) == "HEAD" && back.substr(4, n) == payload &&
                              back.substr(4 + n) == "TAIL"
                          ? "roundtrip-ok"
                          : "roundtrip-BAD")
                  << "\n";
    } else {
        return 2;
    }
    return 0;
}
