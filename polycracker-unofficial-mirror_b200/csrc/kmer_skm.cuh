// Super-k-mer partitioning for K3 (host/device core; the kernels are in kernels_skm.cuh).
//
// The round-1 count moved one 8-byte canonical k-mer per window through two radix scatters before it reached a
// shared-memory table: 28 GB of DRAM traffic for 12 GB of algorithmic bytes, and both scatters bound by per-k-mer
// shared-memory atomics.  Here the unit that travels is a SUPER-K-MER: a run of consecutive windows whose k-mers share
// the same minimizer (the canonical m-mer with the smallest hash among the w = k - m + 1 m-mers of the window; KMC 2 /
// Gerbil style signatures).  The bucket of a k-mer is a function of its minimizer only, the minimizer is a function of
// the k-mer's content and is the same for the k-mer and its reverse complement, so every instance of a canonical k-mer
// lands in the same bucket -- counting bucket by bucket stays exact -- while a run of n windows travels as n + k - 1
// packed bases (2 bits each) instead of n x 64 bits.  The k-mers are expanded again only inside the count kernel, in
// registers, right in front of the shared-memory table.
//
// Record (16 bytes, two uint64 hi:lo):   bits 127..30  up to 49 bases, first base in bits 127:126
//                                        bits  29..6   destination = (level-1 bucket << 11) | sub-bucket
//                                        bits   5..0   number of k-mers - 1
#pragma once
#include "kmer_device.cuh"

namespace pc {

#define PC_SKM_BASES 49                 // bases one record holds -> at most 49 - k + 1 k-mers
#define PC_SKM_SUB_BITS 11              // sub-buckets per level-1 bucket <= 2048 (PC_SUB_MAX)
#define PC_SKM_WMIN 4
#define PC_SKM_WMAX 22

// hash of a canonical m-mer (2m <= 32 bits): a bijection of uint32, so equal hashes <=> equal m-mers and the minimum
// hash identifies ONE m-mer.  The seed keeps poly-A (code 0) from being the global minimum.
// One multiply and one fold: the extraction is integer-ALU bound (ncu: SM 84 %) and hashes 32 + W - 1 m-mers per thread;
// the minimum is decided by the high bits of the product, which depend on every bit of the m-mer, and the destination
// is mixed again in skm_dest (the second multiply round used here first cost 0.1 ms and balanced the buckets no better).
PC_HD uint32_t mm_hash(uint32_t c) {
    c ^= 0x6A09E667u; c *= 0x9E3779B1u; c ^= c >> 15;
    return c;
}

// destination of a minimizer: the minimum of w hashes is biased towards small values, so it is mixed again before its
// top pb bits choose the level-1 bucket and the following bits the sub-bucket (same split as sub_of() for k-mers).
PC_HD uint32_t skm_dest(uint32_t hmin, int pb, uint32_t nb2) {
    uint32_t r = hmin * 0xC2B2AE35u; r ^= r >> 16; r *= 0x27D4EB2Fu; r ^= r >> 15;
    uint32_t q = pb ? (r >> (32 - pb)) : 0u;
    uint32_t rest = pb ? (r << pb) : r;
    rest ^= rest >> 13;                                        // the bits shifted in at the bottom are zero: fold
    uint32_t sub = (uint32_t)(((uint64_t)rest * (uint64_t)nb2) >> 32);
    return (q << PC_SKM_SUB_BITS) | sub;
}

PC_HD uint32_t revcomp32(uint32_t x, int m) {                  // m-mer in the low 2m bits, m <= 16
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    x = (x >> 16) | (x << 16);
    return (~x) >> (32 - 2 * m);
}

// Everything one thread derives from its packed word (32 window starts) and the word after it:
//   vmask  bit b: window b is valid (no masked base in [b, b + k): N, soft padding, chunk end)
//   starts bit b: a record starts at window b (first valid window, minimizer change, or the run reached nmax)
//   hmin[b]     : minimum m-mer hash of window b (meaningful for valid windows)
// W = k - m + 1 is a template parameter so that every array index below is a compile-time constant (registers).
template <int W>
struct SkmThread {
    static constexpr int NP = 32 + W - 1;                      // m-mer positions the 32 windows touch
    uint32_t vmask, starts;
    uint32_t hmin[32];

    PC_HD void run(uint64_t w0, uint64_t w1, uint32_t m0, uint32_t m1, int k, int m, int nmax) {
        // ---- window validity: dilate the masked-base bits by k - 1 towards earlier window starts
        uint64_t r = ((uint64_t)m0 << 32) | (uint64_t)m1;      // bit (63 - j): base j is masked
        int width = 1;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int s = 1; s <= 16; s <<= 1)
            if (2 * s <= k) { r |= r << s; width = 2 * s; }
        r |= r << (k - width);                                 // k - width < width
        const uint32_t vmsb = ~(uint32_t)(r >> 32);            // bit (31 - b): window b valid

        // ---- rolling canonical m-mer hashes at positions 0 .. NP - 1
        const uint32_t mmask = m >= 16 ? 0xFFFFFFFFu : ((1u << (2 * m)) - 1u);
        const int rs = 2 * (m - 1);
        uint32_t fwd = (uint32_t)(w0 >> (64 - 2 * m));
        uint32_t rc = revcomp32(fwd, m);
        // bases m, m + 1, ... at the top of a 128-bit value, held as four 32-bit words (t[0] most significant)
        const uint64_t thi = (w0 << (2 * m)) | (w1 >> (64 - 2 * m));   // 2 <= 2m <= 32
        const uint64_t tlo = w1 << (2 * m);
        const uint32_t t[4] = {(uint32_t)(thi >> 32), (uint32_t)thi, (uint32_t)(tlo >> 32), (uint32_t)tlo};
        uint32_t h[NP];
        h[0] = mm_hash(fwd < rc ? fwd : rc);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 1; j < NP; ++j) {
            const int i = j - 1;                               // incoming base m + i
            const uint32_t nb = (t[i / 16] >> (30 - 2 * (i % 16))) & 3u;
            fwd = ((fwd << 2) | nb) & mmask;
            rc = (rc >> 2) | ((nb ^ 3u) << rs);
            h[j] = mm_hash(fwd < rc ? fwd : rc);
        }
        // ---- sliding minimum over W consecutive positions, in place: doubling, then one combine
        int c = 1;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int lvl = 0; lvl < 5; ++lvl) {
            if (2 * c <= W) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int j = 0; j + 2 * c <= NP; ++j) h[j] = h[j] < h[j + c] ? h[j] : h[j + c];
                c *= 2;
            }
        }
        const int d = W - c;                                   // 0 <= d < c
        // ---- record boundaries
        uint32_t vm = 0, st = 0, prev = 0;
        bool pv = false;
        int runlen = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int b = 0; b < 32; ++b) {
            const uint32_t hm = h[b] < h[b + d] ? h[b] : h[b + d];
            hmin[b] = hm;
            const bool v = (vmsb >> (31 - b)) & 1u;
            const bool s = v && (!pv || hm != prev || runlen >= nmax);
            runlen = s ? 1 : runlen + 1;
            vm |= (uint32_t)v << b;
            st |= (uint32_t)s << b;
            prev = hm; pv = v;
        }
        vmask = vm; starts = st;
    }
};

// number of windows of the record that starts at window b: up to the next record start / invalid window / word end
PC_HD int skm_run_length(uint32_t vmask, uint32_t starts, int b) {
    const uint32_t after = (starts | ~vmask) & (0xFFFFFFFEu << b);
    int e = 32;
    if (after) {
#if defined(__CUDA_ARCH__)
        e = __ffs(after) - 1;
#else
        e = __builtin_ctz(after);
#endif
    }
    return e - b;
}

// the record of n windows starting at window s of (w0, w1)
PC_HD void skm_record(uint64_t w0, uint64_t w1, int s, int n, int k, uint32_t dest, uint64_t& hi, uint64_t& lo) {
    const int sh = 2 * s;                                      // 0 .. 62
    hi = sh ? ((w0 << sh) | (w1 >> (64 - sh))) : w0;
    lo = w1 << sh;
    const int L = n + k - 1;                                   // bases kept, <= PC_SKM_BASES
    if (L <= 32) { if (L < 32) hi &= ~(~0ull >> (2 * L)); lo = 0ull; }
    else lo &= ~(~0ull >> (2 * (L - 32)));                     // 1 <= L - 32 <= 17
    lo |= ((uint64_t)dest << 6) | (uint64_t)(n - 1);
}

PC_HD uint32_t skm_rec_dest(uint64_t lo) { return (uint32_t)(lo >> 6) & 0xFFFFFFu; }
PC_HD int skm_rec_n(uint64_t lo) { return (int)(lo & 63ull) + 1; }

// rolling canonical k-mers of one record
struct RecRoller {
    uint64_t fwd, rc, t, kmask;
    int rcshift;
    PC_HD void init(uint64_t hi, uint64_t lo, int k) {
        kmask = kmer_mask(k);
        rcshift = 2 * (k - 1);
        fwd = hi >> (64 - 2 * k);
        rc = revcomp_bits(fwd, k);
        t = k >= 32 ? lo : ((hi << (2 * k)) | (lo >> (64 - 2 * k)));   // bases k, k + 1, ... at the top
    }
    PC_HD uint64_t canonical() const { return fwd < rc ? fwd : rc; }
    PC_HD void step() {
        const uint64_t nb = t >> 62;
        t <<= 2;
        fwd = ((fwd << 2) | nb) & kmask;
        rc = (rc >> 2) | ((3ull - nb) << rcshift);
    }
};

// canonical k-mer number j of a record, extracted directly (no rolling): the count kernel spreads the k-mers of a warp's
// 32 records evenly over its lanes, so a lane needs k-mer j of somebody else's record
PC_HD uint64_t skm_rec_kmer(uint64_t hi, uint64_t lo, int j, int k) {
    const uint64_t x = j ? ((hi << (2 * j)) | (lo >> (64 - 2 * j))) : hi;      // j <= 31
    const uint64_t fwd = x >> (64 - 2 * k);
    const uint64_t rc = revcomp_bits(fwd, k);
    return fwd < rc ? fwd : rc;
}

// plan shared by every rank: minimizer length for k and the number of sub-buckets.  W = k - m + 1 even, in
// [PC_SKM_WMIN, PC_SKM_WMAX], m in [9, 16]; returns 0 when k is too short for this path.
inline int skm_choose_m(int k, int64_t n_sub_total) {
    int m = 10;
    while (m < 16 && ((int64_t)1 << (2 * m - 1)) < n_sub_total * 64) ++m;   // >= 64 canonical m-mers per sub-bucket
    int W = k - m + 1;
    // W must be even (the sliding minimum doubles).  Rounding m UP keeps the minimizer buckets light: with the shorter m a
    // 2 Gbp genome put 8 406 of 485 k sub-buckets beyond their shared-memory table (fall-back 7.5 ms, count 21.2 ms) against
    // 81 with m + 1 (0.14 ms, 18.6 ms), for 15 % more records through the scatters; at 0.5 Gbp the two are equal.
    if (W & 1) { if (m < 16 && W - 1 >= PC_SKM_WMIN) { ++m; --W; } else { --m; ++W; } }
    while (W < PC_SKM_WMIN && m > 9) { m -= 2; W += 2; }
    while (W > PC_SKM_WMAX) { m += 2; W -= 2; }
    if (W < PC_SKM_WMIN || m < 9 || m > 16 || m > k) return 0;
    return m;
}

}  // namespace pc
