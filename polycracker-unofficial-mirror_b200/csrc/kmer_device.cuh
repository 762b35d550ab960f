// Shared host/device helpers for the polyCRACKER hot path on B200 (sm_100a).
//
// Data layout in HBM (see DESIGN.md "Data layout"):
//   * words[w]  : uint64, 32 bases, base j of the word in bits (63-2j, 62-2j), A=0 C=1 G=2 T=3, so that an
//                 unsigned integer compare of two k-mers equals the Python string compare the reference uses
//                 for its column order (polycracker.py:295, sorted()).
//   * nmask[w]  : uint32, bit (31-j) set when base j is NOT one of acgtACGT, or is chunk padding.
//   * every chunk (= one record of <genome>_split.fa, polycracker.py:170) starts on a word boundary and is
//     followed by >= 1 padding base, so a window that would leave its chunk always covers an invalid base
//     and is rejected by the same test that rejects N's: no k-mer ever spans two chunks
//     (kmercountexact/jellyfish/bbmap treat each FASTA record independently, polycracker.py:199,203,252).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PC_HD __host__ __device__ __forceinline__
#else
#define PC_HD inline
#endif

namespace pc {

// ---- ASCII -> 2-bit, 4 bases at a time (SWAR).  `v` holds 4 ASCII bytes, first base in the low byte. ----
// code = ((c>>1) ^ (c>>2)) & 3 maps A/a->0, C/c->1, G/g->2, T/t->3.
// Returns 8 bits: base0 in bits 7:6 ... base3 in bits 1:0.
PC_HD uint32_t pack4_codes(uint32_t v) {
    uint32_t x = ((v >> 1) ^ (v >> 2)) & 0x03030303u;
    return (x * 0x40100401u) >> 24;
}

// Returns 4 bits (base0 -> bit 3 ... base3 -> bit 0), set when the byte is not one of acgtACGT.
PC_HD uint32_t pack4_invalid(uint32_t v) {
    uint32_t u = v & 0xDFDFDFDFu;                       // fold case
    // per-byte equality against 'A','C','G','T' without carries crossing bytes
    uint32_t ok = 0;
    const uint32_t pats[4] = {0x41414141u, 0x43434343u, 0x47474747u, 0x54545454u};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; ++i) {
        uint32_t d = u ^ pats[i];                        // byte == 0 where equal
        // zero-byte detect: ((d & 0x7f7f7f7f) + 0x7f7f7f7f) | d has bit7 set iff byte != 0
        uint32_t nz = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
        ok |= ~nz & 0x80808080u;
    }
    uint32_t bad = (~ok & 0x80808080u) >> 7;            // 0x01 per invalid byte
    return ((bad * 0x08040201u) >> 24) & 0xFu;
}

// ---- reverse complement of a k-mer held in the low 2k bits (first base most significant) ----
PC_HD uint64_t revcomp_bits(uint64_t x, int k) {
#if defined(__CUDA_ARCH__)
    // BREV reverses all 64 bits: the bases end up in reverse order with the two bits of each base swapped; one
    // adjacent-bit swap repairs that (12 instructions instead of ~40 for the five swap stages below)
    uint64_t y = __brevll(x);
    y = ((y & 0x5555555555555555ull) << 1) | ((y >> 1) & 0x5555555555555555ull);
    return (~y) >> (64 - 2 * k);
#endif
    // reverse the order of the 2-bit groups of the whole 64-bit word, then complement and shift down
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
    x = (x >> 32) | (x << 32);
    return (~x) >> (64 - 2 * k);
}

PC_HD uint64_t kmer_mask(int k) { return k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull); }

// 64-bit finaliser (murmur3 fmix64): all output bits depend on all input bits.
PC_HD uint64_t mix64(uint64_t h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 33;
    return h;
}

#if defined(__CUDACC__)
__device__ __forceinline__ uint64_t slot_of(uint64_t h, uint64_t capacity) { return __umul64hi(h, capacity); }
#endif
PC_HD uint64_t slot_of_host(uint64_t h, uint64_t capacity) {
    return (uint64_t)(((unsigned __int128)h * capacity) >> 64);
}

// owner rank of a k-mer for the multi-GPU exchange: a second, independent hash
PC_HD uint32_t owner_of(uint64_t key, uint32_t world) {
    uint64_t h = mix64(key ^ 0x9E3779B97F4A7C15ull);
    return (uint32_t)(((h >> 32) * (uint64_t)world) >> 32);
}

// Rolling canonical k-mer state over one packed word (32 window starts) and its successor.
struct Roller {
    uint64_t fwd, rc, kmask;
    uint64_t w0, w1, m;      // m = (nmask0 << 32) | nmask1
    int k, rcshift;
    PC_HD void init(uint64_t w0_, uint64_t w1_, uint32_t m0, uint32_t m1, int k_) {
        w0 = w0_; w1 = w1_; k = k_;
        m = ((uint64_t)m0 << 32) | (uint64_t)m1;
        kmask = kmer_mask(k);
        rcshift = 2 * (k - 1);
        fwd = w0 >> (64 - 2 * k);
        rc = revcomp_bits(fwd, k);
    }
    // window starting at base b (0..31) of the word is valid iff no invalid base in [b, b+k)
    PC_HD bool valid(int b) const { return ((m << b) >> (64 - k)) == 0ull; }
    PC_HD uint64_t canonical() const { return fwd < rc ? fwd : rc; }
    // advance from window b to window b+1 (b in 0..30)
    PC_HD void step(int b) {
        int j = b + k;                                            // index of the incoming base, 1..63
        uint64_t nb = (j < 32) ? (w0 >> (62 - 2 * j)) & 3ull : (w1 >> (62 - 2 * (j - 32))) & 3ull;
        fwd = ((fwd << 2) | nb) & kmask;
        rc = (rc >> 2) | ((3ull - nb) << rcshift);
    }
};

}  // namespace pc
