// CPU test of the host-side sequence packer (miekki_b200/csrc/pack.cpp): every word is either
// packed exactly (forward digits A0 C1 G2 T3, first base in the top bits) or listed as an exception
// with its raw bytes; both the AVX2 kernel and the portable one.  Prints "ok <backend> <GB/s>".
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <string>
#include <vector>

#include "../../miekki_b200/csrc/pack.h"

static bool check(const std::string& s, int k, uint64_t w0, uint64_t w1) {
    std::vector<uint32_t> out(w1 - w0 + 1, 0xDEADBEEF);
    std::vector<mk::PackException> exc;
    mk::pack_words(s.data(), s.size(), w0, w1, k, out.data(), exc);
    std::map<uint64_t, const mk::PackException*> ex;
    for (auto& e : exc) {
        if (ex.count(e.word)) return false;
        ex[e.word] = &e;
    }
    const uint64_t n = s.size();
    for (uint64_t w = w0; w < w1; ++w) {
        bool plain = 16 * w >= (uint64_t)(k > 1 ? k - 1 : 0) && 16 * w + 16 <= n;
        uint32_t want = 0;
        for (int i = 0; i < 16 && plain; ++i) {
            const char c = s[16 * w + i];
            const char* p = strchr("ACGT", c);
            if (!p || !c) plain = false;
            else want |= (uint32_t)(p - "ACGT") << (30 - 2 * i);
        }
        if (plain) {
            if (ex.count(w) || out[w - w0] != want) return false;
        } else if (16 * w < n) {
            if (!ex.count(w) || out[w - w0] != 0) return false;
            for (int i = 0; i < 16; ++i) {
                const char want_c = 16 * w + i < n ? s[16 * w + i] : 0;
                if ((char)ex[w]->bytes[i] != want_c) return false;
            }
        } else if (ex.count(w) || out[w - w0] != 0) {
            return false;
        }
    }
    return out[w1 - w0] == 0xDEADBEEF && ex.size() == exc.size();
}

int main() {
    std::mt19937_64 rng(7);
    const char alpha[] = "ACGTACGTACGTACGTNacgtRn";
    for (int t = 0; t < 400; ++t) {
        const uint64_t n = rng() % 5000;
        std::string s(n, 'A');
        const bool dirty = t % 3 == 0;
        for (auto& c : s) c = dirty ? alpha[rng() % (sizeof(alpha) - 1)] : "ACGT"[rng() % 4];
        if (t % 5 == 0 && n > 40) s[rng() % n] = 'N';
        const int k = 2 + (int)(rng() % 30);
        const uint64_t nw = (n + 15) / 16;
        if (!check(s, k, 0, nw + 1)) { printf("FAIL whole t=%d n=%llu k=%d\n", t, (unsigned long long)n, k); return 1; }
        if (nw > 3) {
            const uint64_t a = rng() % nw, b = a + rng() % (nw - a + 1);
            if (!check(s, k, a, b)) { printf("FAIL range t=%d\n", t); return 1; }
        }
        if (mk::prefix_is_acgt(s.data(), n, k) !=
            [&] { for (uint64_t i = 0; i < std::min<uint64_t>(n, k - 1); ++i) if (!strchr("ACGTacgt", s[i]) || !s[i]) return false; return true; }()) {
            printf("FAIL prefix t=%d\n", t);
            return 1;
        }
    }
    // throughput of one thread on a 64 Mbp sequence
    std::string big(64u << 20, 'A');
    for (auto& c : big) c = "ACGT"[rng() % 4];
    std::vector<uint32_t> out(big.size() / 16 + 1);
    std::vector<mk::PackException> exc;
    const auto t0 = std::chrono::steady_clock::now();
    for (int rep = 0; rep < 4; ++rep) mk::pack_words(big.data(), big.size(), 0, big.size() / 16, 31, out.data(), exc);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("ok %s %.2f\n", mk::pack_backend(), 4.0 * big.size() / dt / 1e9);
    return 0;
}
