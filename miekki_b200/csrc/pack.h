// pack.h -- host-side 2-bit packing of genome sequences for mk_index_add (host C++, no CUDA).
//
// The sketch kernels read sequences as two 2-bit planes (common.cuh); for a run of 16 upper-case
// A/C/G/T outside the k-1 prefix both planes follow from the forward digits alone.  The host
// therefore ships one 32-bit word per 16 bases (0.25 B/base over PCIe instead of 1 B/base) and,
// for every word that is not such a run -- a word touching the prefix, the ragged last word, a
// word with N / lower case / any other byte -- the 16 raw bytes, which the device encodes with
// the reference's exact rules (utils.cpp:31-49, 107-125, 252-272).
#pragma once
#include <stdint.h>

#include <vector>

namespace mk {

struct PackException {
    uint64_t word;          // index of the word inside the sequence
    uint8_t bytes[16];      // its raw characters, zero padded past the end of the sequence
};

// Packs words [w0, w1) of sequence s (n characters) into out[0 .. w1 - w0): forward digits A0 C1
// G2 T3, base j of a word at bits 30 - 2 j.  Words that need the device's general encoder are
// appended to `exc` (their packed value is 0).  Thread safe for disjoint ranges.
void pack_words(const char* s, uint64_t n, uint64_t w0, uint64_t w1, int k, uint32_t* out,
                std::vector<PackException>& exc);

// str2numstrand's all-or-nothing rule for the first min(k - 1, n) characters (utils.cpp:252-272)
bool prefix_is_acgt(const char* s, uint64_t n, int k);

// "avx2" or "scalar": which packer this machine runs (reported by the bench)
const char* pack_backend();

}  // namespace mk
