// Hand-written sm_100a kernels of the polyCRACKER hot path (K1..K7 of SURVEY.md section 7).
// Everything here is HBM / L2 / atomic bound integer work: no tensor cores, no library calls.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kmer_device.cuh"

namespace pc {

// Hash of a k-mer for bucket / slot selection.  g_hash_mode (set by the host, tunable "hash_mode"):
//   0  murmur3 finaliser (two 64-bit multiplies, three xor-shifts: ~30 integer instructions)
//   1  one 64-bit Fibonacci multiply + xor-fold (~8 instructions); the top bits choose the bucket, the following
//      bits the slot.  Any mode gives the same counts; only probe lengths and bucket balance can differ.
__constant__ int g_hash_mode = 0;
__device__ __forceinline__ uint64_t hash64(uint64_t key) {
    if (g_hash_mode == 1) {
        uint64_t h = key * 0x9E3779B97F4A7C15ull;
        return h ^ (h >> 29);
    }
    return mix64(key);
}
// repeat-table slot (K5).  Deliberately NOT the bucket hash: an owner rank selects k-mers whose bucket hash lies in its
// own 1/world slice of the hash space; reusing those bits would pile all of them into 1/world of the repeat table
// (measured on 4 GPUs: 194 ms for a 750 k-key build instead of 0.03 ms).  rcap < 2^32.
// One 64-bit multiply: the probe kernels are instruction-issue bound once the filters keep most windows away from L2
// (ncu: 63 % of the issue slots busy, 135 warp instructions per 32 windows), and the murmur finaliser used here first
// was a quarter of them.  The top 32 bits of the product depend on every bit of the k-mer.
// (even home slots -- the second step of a chain in the sector the first read brought into L1 -- measured slower: 3.26 vs 3.10 ms)
__device__ __forceinline__ uint64_t rep_slot(uint64_t key, uint64_t rcap) {
    return (((key * 0xD6E8FEB86659FD93ull) >> 32) * rcap) >> 32;
}

// bit of a k-mer in the probe's bitmap filters (shared-memory or L2 resident): an independent multiplicative hash
// (three 32-bit multiplies: the 64-bit product used here first was 6 % of the roller probe's instructions, ncu source page;
// a filter only needs an even spread, and every bit of the k-mer still reaches the top bits of h)
__device__ __forceinline__ unsigned int gbit_of(uint64_t key, unsigned int nbits) {
    const unsigned int h = ((unsigned int)key ^ ((unsigned int)(key >> 32) * 0x85EBCA77u)) * 0x9E3779B1u;
    return __umulhi(h ^ (h >> 15), nbits);
}

// 16-bit fingerprint of a k-mer (never 0), from a third independent multiplicative hash
__device__ __forceinline__ unsigned short fp_of(uint64_t key) {
    return (unsigned short)(((key * 0xC2B2AE3D27D4EB4Full) >> 48) | 1ull);
}

// ------------------------------------------------------------------------------------------------ tables
// Count table (K3): open addressing, linear probing, structure of arrays.  keys[s] = ~kmer so that a zeroed table is
// empty (the all-ones pattern is never a canonical-min k-mer: its reverse complement, all A, is smaller);
// counts[s] = occurrences - 1: the CAS that claims a slot already records the first occurrence, so a new k-mer costs
// one L2 operation and only repeats pay for a RED.  12 bytes per slot; four keys share a 32-byte sector, the threshold
// scan (K4) reads only the 4-byte counts.
#define PC_SLOT_BYTES 12
struct CountTable {
    unsigned long long* keys;
    unsigned int* counts;
    __device__ __forceinline__ unsigned long long* key(uint64_t s) const { return keys + s; }
    __device__ __forceinline__ unsigned int* cnt(uint64_t s) const { return counts + s; }
    __device__ __forceinline__ CountTable sub(uint64_t off) const { return CountTable{keys + off, counts + off}; }
};
// Repeat table slot (K5): selected k-mer -> column id.
struct __align__(16) RepSlot {
    unsigned long long skey;
    int col;
    int pad;
};

struct CountStats {
    unsigned long long valid_windows;
    unsigned long long distinct;
    unsigned long long overflow;      // != 0 when a probe sequence hit the limit (table too small)
    unsigned long long max_probe;
};

// Threshold watch (K4 fused into K3): when low != 0 the increment of a repeat is a returning atomic and the insert
// that lifts a k-mer to exactly `low` occurrences appends it to `keys` -- each k-mer with count >= low exactly once --
// so pc_select(low) does not have to scan the whole table again.
struct Watch {
    unsigned int low;
    unsigned long long* n;
    uint64_t* keys;
    unsigned long long cap;
};
__device__ __forceinline__ void watch_append(const Watch& w, unsigned long long skey) {
    unsigned long long pos = atomicAdd(w.n, 1ull);
    if (pos < w.cap) w.keys[pos] = ~skey;
}
__device__ __forceinline__ void count_repeat(unsigned int* cnt, unsigned long long skey, const Watch& w) {
    if (w.low) {
        unsigned int old = atomicAdd(cnt, 1u);               // stored = occurrences - 1
        if (old + 2u == w.low) watch_append(w, skey);
    } else {
        atomicAdd(cnt, 1u);                                   // result unused -> RED
    }
}

__device__ __forceinline__ unsigned long long ldcg_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ldcg_u4(const void* p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ldnc_u4(const void* p) {   // streaming read-only, do not pollute L1
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// largest c in [0, n) with a[c] <= x   (a ascending, a[0] <= x)
__device__ __forceinline__ int64_t find_le(const int64_t* __restrict__ a, int64_t n, int64_t x) {
    int64_t lo = 0, hi = n;            // invariant a[lo] <= x < a[hi] (a[n] = +inf)
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) <= x) lo = mid; else hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------------ K1 pack
// One thread = one packed word = 32 bases.  A warp reads 1 KiB of contiguous ASCII with 128-bit loads
// (three aligned uint4 per lane, funnel-shifted to the chunk's byte phase) and writes 256 B of codes + 128 B
// of mask, fully coalesced.  chunk_word0[c] is the first word of chunk c; chunk_word0[n_chunks] the total.
// ascii_lo/ascii_hi bound the bytes that may be touched by the aligned over-read (the staging buffer is padded).
template <int WS>
__device__ __forceinline__ void realign8(const uint32_t (&r)[12], int bs, uint32_t (&o)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = __funnelshift_r(r[WS + i], r[WS + i + 1], bs);
}

// TMA (1-D bulk copy) helpers: one elected thread arms an mbarrier with the byte count and issues
// cp.async.bulk global -> shared; everybody waits on the barrier phase.  SASS: UBLKCP / SYNCS.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned int bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// One CTA = 256 words = 8 KiB of ASCII.  When the CTA's words lie in one chunk (the normal case: chunks are tens of
// kilobases) its ASCII range is staged in shared memory with ONE bulk copy (TMA engine, fully coalesced 16-byte
// aligned HBM reads, no LSU instructions) and every thread converts its 32 bases out of shared memory; CTAs that
// straddle a chunk boundary or touch the end of a caller-owned buffer use per-thread 128-bit global loads.
#define PC_PACK_TILE_BYTES 8192
__global__ void __launch_bounds__(256)
k_pack(const uint8_t* __restrict__ ascii, int64_t ascii_bytes_padded,
       const int64_t* __restrict__ chunk_word0, const int64_t* __restrict__ chunk_begin,
       const int64_t* __restrict__ chunk_len, int64_t n_chunks, int64_t n_words_total,
       uint64_t* __restrict__ words, uint32_t* __restrict__ nmask)
{
    __shared__ __align__(128) uint8_t stage[PC_PACK_TILE_BYTES + 64];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ long long s_c0, s_a0;
    __shared__ int s_bulk;
    const int64_t data_words = __ldg(chunk_word0 + n_chunks);
    const int64_t w_first = (int64_t)blockIdx.x * blockDim.x;
    if (threadIdx.x == 0) {
        int bulk = 0; long long c0 = 0, a0 = 0;
        if (n_chunks > 0 && w_first < data_words) {
            c0 = find_le(chunk_word0, n_chunks, w_first);
            int64_t w_last = min(w_first + (int64_t)blockDim.x, data_words) - 1;
            int64_t off = (w_first - __ldg(chunk_word0 + c0)) * 32;
            int64_t avail = __ldg(chunk_len + c0) - off;                     // bases of the chunk from this CTA on
            if (w_last < __ldg(chunk_word0 + c0 + 1) && avail > 0) {
                int64_t src = __ldg(chunk_begin + c0) + off;
                int sh = (int)((reinterpret_cast<uintptr_t>(ascii) + (uintptr_t)src) & 15);
                a0 = src - sh;
                int64_t need = min((int64_t)PC_PACK_TILE_BYTES, avail);
                int64_t size = (sh + need + 15) & ~(int64_t)15;               // <= 8192 + 16
                // the per-thread realignment reads up to 48 bytes from its 16-byte aligned start: keep that inside
                // the staged range by copying 48 more bytes when the buffer allows it
                size = min(size + 48, (int64_t)(PC_PACK_TILE_BYTES + 64));
                if (a0 >= 0 && a0 + size <= ascii_bytes_padded) {
                    bulk = 1;
                    mbar_init(&bar, 1);
                    mbar_expect_tx(&bar, (unsigned int)size);
                    bulk_g2s(stage, ascii + a0, (unsigned int)size, &bar);
                }
            }
        }
        s_bulk = bulk; s_c0 = c0; s_a0 = a0;
    }
    __syncthreads();
    const bool bulk = s_bulk != 0;
    if (bulk) mbar_wait(&bar, 0);
    int64_t w = w_first + threadIdx.x;
    if (w >= n_words_total) return;
    if (w >= data_words) {             // trailing guard word(s): all padding
        words[w] = 0ull; nmask[w] = 0xFFFFFFFFu; return;
    }
    int64_t c = s_c0;                  // chunk of the CTA's first word; walk forward (rarely more than one step)
    while (c + 1 < n_chunks && __ldg(chunk_word0 + c + 1) <= w) ++c;
    int64_t off = (w - __ldg(chunk_word0 + c)) * 32;
    int64_t remaining = __ldg(chunk_len + c) - off;         // may be <= 0: pure padding word
    uint64_t word = 0ull; uint32_t nm = 0xFFFFFFFFu;
    if (remaining > 0) {
        int64_t src = __ldg(chunk_begin + c) + off;          // byte offset into ascii
        int sh = (int)((reinterpret_cast<uintptr_t>(ascii) + (uintptr_t)src) & 15);
        int64_t a0 = src - sh;                                // 16-byte aligned
        uint32_t o[8];
        bool have = false;
        uint4 q0, q1, q2;
        if (bulk) {
            const uint8_t* p = stage + (a0 - s_a0);           // 16-byte aligned inside the staged range
            q0 = *reinterpret_cast<const uint4*>(p);
            q1 = *reinterpret_cast<const uint4*>(p + 16);
            q2 = *reinterpret_cast<const uint4*>(p + 32);
            have = true;
        } else if (a0 >= 0 && a0 + 48 <= ascii_bytes_padded) {
            const uint8_t* p = ascii + a0;
            q0 = ldnc_u4(p); q1 = ldnc_u4(p + 16); q2 = ldnc_u4(p + 32);
            have = true;
        }
        if (have) {
            uint32_t r[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
            int bs = (sh & 3) * 8;
            switch (sh >> 2) {
                case 0: realign8<0>(r, bs, o); break;
                case 1: realign8<1>(r, bs, o); break;
                case 2: realign8<2>(r, bs, o); break;
                default: realign8<3>(r, bs, o); break;
            }
        } else {                                              // caller-owned device buffer without slack
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t v = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    int64_t idx = src + 4 * i + b;
                    uint32_t ch = (4 * i + b < remaining && idx < ascii_bytes_padded) ? ascii[idx] : (uint32_t)'N';
                    v |= ch << (8 * b);
                }
                o[i] = v;
            }
        }
        nm = 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            word |= (uint64_t)pack4_codes(o[i]) << (56 - 8 * i);
            nm |= pack4_invalid(o[i]) << (28 - 4 * i);
        }
        if (remaining < 32) {                                 // bases past the chunk end are padding
            nm |= 0xFFFFFFFFu >> (int)remaining;
            word &= ~(~0ull >> (2 * (int)remaining));
        }
        // invalid bases carry code 0 so the packed words are deterministic
        // (spread the 32 mask bits to 64 code bits: bit (31-j) -> bits (63-2j, 62-2j))
        uint64_t m64 = nm;
        m64 = (m64 | (m64 << 16)) & 0x0000FFFF0000FFFFull;
        m64 = (m64 | (m64 << 8)) & 0x00FF00FF00FF00FFull;
        m64 = (m64 | (m64 << 4)) & 0x0F0F0F0F0F0F0F0Full;
        m64 = (m64 | (m64 << 2)) & 0x3333333333333333ull;
        m64 = (m64 | (m64 << 1)) & 0x5555555555555555ull;
        word &= ~(m64 | (m64 << 1));
    }
    words[w] = word;
    nmask[w] = nm;
}

// ------------------------------------------------------------------------------------------------ K2 extraction
// One lane = one packed word = 32 window starts; the (k-1)-base overlap comes from the next lane's word by
// warp shuffle (lane 31 loads it), so every packed word is read from HBM once.
struct WordPair {
    uint64_t w0, w1; uint32_t m0, m1;
};
__device__ __forceinline__ WordPair load_word_pair(const uint64_t* __restrict__ words,
                                                   const uint32_t* __restrict__ nmask,
                                                   int64_t w, int64_t n_words /* data words; word n_words is a guard */)
{
    WordPair p;
    bool in = w < n_words;
    p.w0 = in ? __ldg(words + w) : 0ull;
    p.m0 = in ? __ldg(nmask + w) : 0xFFFFFFFFu;
    p.w1 = __shfl_down_sync(0xFFFFFFFFu, p.w0, 1);
    p.m1 = __shfl_down_sync(0xFFFFFFFFu, p.m0, 1);
    if ((threadIdx.x & 31) == 31) {
        bool in1 = (w + 1) < n_words;
        p.w1 = in1 ? __ldg(words + w + 1) : 0ull;
        p.m1 = in1 ? __ldg(nmask + w + 1) : 0xFFFFFFFFu;
    }
    return p;
}

// ------------------------------------------------------------------------------------------------ K3 count
#define PC_PROBE_LIMIT (1u << 22)

__device__ __forceinline__ void count_insert(CountTable table, uint64_t capacity, uint64_t slot,
                                             unsigned long long skey, unsigned long long cur,
                                             unsigned int& new_keys, unsigned int& max_probe,
                                             CountStats* stats, const Watch w = Watch{0u, nullptr, nullptr, 0ull})
{
    // `cur` is the (possibly stale-empty) key already loaded from table[slot]
    unsigned int probes = 0;
    for (;;) {
        if (cur == 0ull) {
            cur = atomicCAS(table.key(slot), 0ull, skey);
            if (cur == 0ull) {                                 // claimed: count field 0 == one occurrence
                ++new_keys;
                if (w.low == 1u) watch_append(w, skey);
                break;
            }
        }
        if (cur == skey) { count_repeat(table.cnt(slot), skey, w); break; }
        if (++probes > PC_PROBE_LIMIT || probes >= capacity) { atomicAdd(&stats->overflow, 1ull); break; }
        if (++slot == capacity) slot = 0;
        cur = ldcg_u64(table.key(slot));
    }
    max_probe = max(max_probe, probes);
}

// CAS-first variant for L2-resident sub-tables: no separate load, one ATOM per new k-mer, ATOM + RED per repeat.
// `cur` is the value the first CAS on table[slot] returned.
__device__ __forceinline__ void count_insert_cas(CountTable table, uint64_t capacity, uint64_t slot,
                                                 unsigned long long skey, unsigned long long cur,
                                                 unsigned int& new_keys, unsigned int& max_probe, CountStats* stats,
                                                 const Watch w = Watch{0u, nullptr, nullptr, 0ull})
{
    unsigned int probes = 0;
    for (;;) {
        if (cur == 0ull) { ++new_keys; if (w.low == 1u) watch_append(w, skey); break; }
        if (cur == skey) { count_repeat(table.cnt(slot), skey, w); break; }
        if (++probes > PC_PROBE_LIMIT || probes >= capacity) { atomicAdd(&stats->overflow, 1ull); break; }
        if (++slot == capacity) slot = 0;
        cur = atomicCAS(table.key(slot), 0ull, skey);
    }
    max_probe = max(max_probe, probes);
}

// insert phases shared by the count kernels: the table loads of a batch are already in flight / landed in cur[];
// (1) every CAS the batch needs is issued back to back, (2) results are folded, (3) REDs go out, and only keys that
// collided with a different k-mer enter the (serial) linear-probing loop.
template <int N>
__device__ __forceinline__ void count_insert_batch(CountTable table, uint64_t capacity, const uint64_t (&slot)[N],
                                                   const unsigned long long (&skey)[N], unsigned long long (&cur)[N],
                                                   const bool (&on)[N], unsigned int& new_keys, unsigned int& max_probe,
                                                   CountStats* stats, const Watch w = Watch{0u, nullptr, nullptr, 0ull})
{
    bool cas[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        cas[j] = on[j] && cur[j] == 0ull;
        if (cas[j]) cur[j] = atomicCAS(table.key(slot[j]), 0ull, skey[j]);
    }
    bool claimed[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        claimed[j] = cas[j] && cur[j] == 0ull;
        new_keys += claimed[j];
        if (claimed[j] && w.low == 1u) watch_append(w, skey[j]);
    }
    if (w.low) {                                               // returning atomics, all issued before any is consumed
        unsigned int old[N];
#pragma unroll
        for (int j = 0; j < N; ++j)
            old[j] = (on[j] && !claimed[j] && cur[j] == skey[j]) ? atomicAdd(table.cnt(slot[j]), 1u) : 0xFFFFFFF0u;
#pragma unroll
        for (int j = 0; j < N; ++j)
            if (old[j] + 2u == w.low) watch_append(w, skey[j]);
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j)
            if (on[j] && !claimed[j] && cur[j] == skey[j]) atomicAdd(table.cnt(slot[j]), 1u);
    }
#pragma unroll
    for (int j = 0; j < N; ++j)
        if (on[j] && !claimed[j] && cur[j] != skey[j]) {       // occupied by another k-mer: probe onwards
            uint64_t s = slot[j] + 1; if (s == capacity) s = 0;
            unsigned long long c2 = ldcg_u64(table.key(s));
            unsigned int mp = 0;
            count_insert(table, capacity, s, skey[j], c2, new_keys, mp, stats, w);
            max_probe = max(max_probe, mp + 1u);
        }
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}

// MODE 0: load the slot key, then CAS/RED.   MODE 1: as 0, but a second roller runs one batch ahead and issues
// prefetch.global.L2 for the slots of the next batch (memory-level parallelism without holding registers).
// MODE 2: CAS first (no separate load): one ATOM + one RED per insert.
template <int BATCH, int MINB, int MODE>
__global__ void __launch_bounds__(256, MINB)
k_count(const uint64_t* __restrict__ words, const uint32_t* __restrict__ nmask, int64_t n_words, int k,
        CountTable table, uint64_t capacity, CountStats* __restrict__ stats)
{
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    WordPair p = load_word_pair(words, nmask, w, n_words);
    unsigned int nvalid = 0, new_keys = 0, max_probe = 0;
    if (p.m0 != 0xFFFFFFFFu) {
        Roller r; r.init(p.w0, p.w1, p.m0, p.m1, k);
        Roller lead = r;
        if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < BATCH; ++j) {
                if (lead.valid(j)) prefetch_l2(table.key(slot_of(hash64(lead.canonical()), capacity)));
                lead.step(j);
            }
        }
#pragma unroll 1
        for (int b0 = 0; b0 < 32; b0 += BATCH) {
            if (MODE == 1 && b0 + BATCH < 32) {
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    int b = b0 + BATCH + j;
                    if (lead.valid(b)) prefetch_l2(table.key(slot_of(hash64(lead.canonical()), capacity)));
                    if (b < 31) lead.step(b);
                }
            }
            unsigned long long skey[BATCH]; uint64_t slot[BATCH]; unsigned long long cur[BATCH];
            unsigned int vmask = 0;
#pragma unroll
            for (int j = 0; j < BATCH; ++j) {
                int b = b0 + j;
                uint64_t key = r.canonical();
                bool v = r.valid(b);
                vmask |= (unsigned int)v << j;
                skey[j] = ~key;
                slot[j] = slot_of(hash64(key), capacity);
                if (b < 31) r.step(b);
            }
#pragma unroll
            for (int j = 0; j < BATCH; ++j) {                  // all probes of the batch in flight together
                if (MODE == 2) {
                    cur[j] = ((vmask >> j) & 1u) ? atomicCAS(table.key(slot[j]), 0ull, skey[j]) : 0ull;
                } else {
                    cur[j] = ((vmask >> j) & 1u) ? ldcg_u64(table.key(slot[j])) : 0ull;
                }
            }
            if (MODE == 2) {
#pragma unroll
                for (int j = 0; j < BATCH; ++j)
                    if ((vmask >> j) & 1u)
                        count_insert_cas(table, capacity, slot[j], skey[j], cur[j], new_keys, max_probe, stats);
            } else {
                bool on[BATCH];
#pragma unroll
                for (int j = 0; j < BATCH; ++j) on[j] = (vmask >> j) & 1u;
                count_insert_batch<BATCH>(table, capacity, slot, skey, cur, on, new_keys, max_probe, stats);
            }
            nvalid += __popc(vmask);
        }
    }
    // per-warp statistics: 2-3 atomics per 1024 windows
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nvalid += __shfl_xor_sync(0xFFFFFFFFu, nvalid, o);
        new_keys += __shfl_xor_sync(0xFFFFFFFFu, new_keys, o);
        max_probe = max(max_probe, __shfl_xor_sync(0xFFFFFFFFu, max_probe, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (nvalid) atomicAdd(&stats->valid_windows, (unsigned long long)nvalid);
        if (new_keys) atomicAdd(&stats->distinct, (unsigned long long)new_keys);
        if (max_probe > 8) atomicMax(&stats->max_probe, (unsigned long long)max_probe);
    }
}

// ------------------------------------------------------------------------------------------------ K3, partitioned
// Random inserts into a table much larger than L2 are bound by DRAM row activations: every insert fetches a
// whole 128-byte line and later writes a sector back (ncu: 150 B of DRAM traffic per insert, 26 % of peak
// bandwidth, 30 us queueing latency).  The partitioned path restores locality: the table is P = 2^pb independent
// sub-tables of S slots (partition = top pb bits of the hash), the k-mers are first radix-partitioned into P
// buckets with streaming writes, then the buckets are counted one after the other, so the sub-table being
// updated stays L2 resident and every table line moves between DRAM and L2 once.
__device__ __forceinline__ unsigned int part_of(uint64_t h, int pb) { return pb ? (unsigned int)(h >> (64 - pb)) : 0u; }
__device__ __forceinline__ uint64_t slot_in_part(uint64_t h, int pb, uint64_t S) { return __umul64hi(h << pb, S); }

#define PC_KEY_SENTINEL (~0ull)      // never a canonical-min k-mer (its reverse complement, all A, is smaller)

// pass A: packed genome -> stream of canonical k-mers, one slot per window start (invalid windows hold the
// sentinel), plus the global bucket histogram.  Stream layout [warp tile][window b][lane]: every store
// instruction writes 256 contiguous bytes.
__global__ void __launch_bounds__(256)
k_extract(const uint64_t* __restrict__ words, const uint32_t* __restrict__ nmask, int64_t n_words, int k, int pb,
          uint64_t* __restrict__ stream, unsigned long long* __restrict__ bucket_hist, CountStats* __restrict__ stats)
{
    extern __shared__ unsigned int shh[];                     // [P] bucket counts of this CTA
    const unsigned int P = 1u << pb;
    for (unsigned int i = threadIdx.x; i < P; i += blockDim.x) shh[i] = 0;
    __syncthreads();
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int lane = threadIdx.x & 31;
    WordPair p = load_word_pair(words, nmask, w, n_words);
    uint64_t* out = stream + (w - lane) * 32 + lane;
    unsigned int nvalid = 0;
    if (p.m0 != 0xFFFFFFFFu) {
        Roller r; r.init(p.w0, p.w1, p.m0, p.m1, k);
#pragma unroll 8
        for (int b = 0; b < 32; ++b) {
            bool v = r.valid(b);
            uint64_t key = r.canonical();
            __stcs(out + b * 32, v ? key : PC_KEY_SENTINEL);
            if (v) { atomicAdd(&shh[part_of(hash64(key), pb)], 1u); ++nvalid; }
            if (b < 31) r.step(b);
        }
    } else {
#pragma unroll 8
        for (int b = 0; b < 32; ++b) __stcs(out + b * 32, PC_KEY_SENTINEL);
    }
    __syncthreads();
    for (unsigned int i = threadIdx.x; i < P; i += blockDim.x)
        if (shh[i]) atomicAdd(&bucket_hist[i], (unsigned long long)shh[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(0xFFFFFFFFu, nvalid, o);
    if (lane == 0 && nvalid) atomicAdd(&stats->valid_windows, (unsigned long long)nvalid);
}

// exclusive scan of the P bucket sizes -> bucket_off[P+1]; cursors reset.  P <= 4096: one CTA.
__global__ void __launch_bounds__(1024)
k_bucket_scan(const unsigned long long* __restrict__ hist, unsigned int P, unsigned long long* __restrict__ bucket_off,
              unsigned long long* __restrict__ cursor)
{
    __shared__ unsigned long long part[1024];
    int t = threadIdx.x;
    unsigned int per = (P + 1023) / 1024;
    unsigned int beg = min(P, t * per), end = min(P, beg + per);
    unsigned long long s = 0;
    for (unsigned int i = beg; i < end; ++i) s += hist[i];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        unsigned long long v = (t >= o) ? part[t - o] : 0ull;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    unsigned long long run = part[t] - s;
    for (unsigned int i = beg; i < end; ++i) { bucket_off[i] = run; cursor[i] = 0ull; run += hist[i]; }
    if (t == 1023) bucket_off[P] = part[1023];
}

// pass B: tile-wise counting sort by bucket in shared memory, then coalesced copy-out.  A tile (TILE stream slots)
// is histogrammed, one global atomic per non-empty (tile, bucket) reserves a contiguous run in the bucket, the
// keys are ranked into a bucket-sorted smem array and written out run by run: consecutive threads write
// consecutive addresses (scattered 8-byte stores into hundreds of open runs were the bottleneck of the first
// version: 42 G stores/s regardless of how the ranks were computed).
// PACK (k <= 26: the k-mer leaves the top 12 bits of its word free): the bucket id travels in those bits of the sorted
// shared array instead of a second array, and gbase holds (global run base - local run start), so the copy-out costs two
// shared loads per k-mer instead of four and the ranking pass one shared store instead of two.  These kernels are bound
// by the shared-memory pipe (ncu: L1/shared 65-71 % busy at 25 % occupancy; more resident CTAs changed nothing).
#define PC_PACK_SHIFT 52
template <int TILE, bool PACK>
__global__ void __launch_bounds__(256, TILE <= 4096 ? 4 : 2)
k_tile_scatter(const uint64_t* __restrict__ stream, int64_t n_stream, int pb,
               const unsigned long long* __restrict__ bucket_off, unsigned long long* __restrict__ cursor,
               uint64_t* __restrict__ kbuf)
{
    constexpr int PER = TILE / 256;                                      // keys per thread, held in registers
    extern __shared__ unsigned long long sm64[];
    const unsigned int P = 1u << pb;
    unsigned long long* sorted = sm64;                                   // [TILE]
    unsigned long long* gbase = sm64 + TILE;                             // [P]
    unsigned int* cnt = reinterpret_cast<unsigned int*>(gbase + P);      // [P]
    unsigned int* lstart = cnt + P;                                      // [P + 1]
    unsigned short* qs = reinterpret_cast<unsigned short*>(lstart + P + 1 + ((P + 1) & 1));   // [TILE]
    __shared__ unsigned int wsum[8];
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    int64_t n_tiles = (n_stream + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t* src = stream + tile * TILE;
        int64_t lim = min((int64_t)TILE, n_stream - tile * TILE);
        // all PER loads of a thread are issued before any is consumed (the first version interleaved each load with
        // its smem atomic and sat on long-scoreboard stalls at 25 % occupancy); the keys stay in registers for the
        // ranking pass, so the stream is read once
        uint64_t key[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            int i = j * 256 + t;
            key[j] = (i < lim) ? __ldcs(src + i) : PC_KEY_SENTINEL;
        }
        for (unsigned int i = t; i < P; i += 256) cnt[i] = 0;
        __syncthreads();
        unsigned short q[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            q[j] = (unsigned short)part_of(hash64(key[j]), pb);
            if (key[j] != PC_KEY_SENTINEL) atomicAdd(&cnt[q[j]], 1u);
        }
        __syncthreads();
        // exclusive scan of cnt[P] -> lstart; reserve the global runs; cnt becomes the rank counter
        {
            unsigned int per = (P + 255) / 256;
            unsigned int b0 = min(P, t * per), b1 = min(P, b0 + per);
            unsigned int s = 0;
            for (unsigned int x = b0; x < b1; ++x) s += cnt[x];
            unsigned int incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned int u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += u;
            }
            if (lane == 31) wsum[wid] = incl;
            __syncthreads();
            unsigned int before = 0;
#pragma unroll
            for (int x = 0; x < 8; ++x) if (x < wid) before += wsum[x];
            unsigned int run = before + incl - s;
            for (unsigned int x = b0; x < b1; ++x) {
                unsigned int c = cnt[x];
                lstart[x] = run;
                if (c) gbase[x] = bucket_off[x] + atomicAdd(&cursor[x], (unsigned long long)c) - run;   // minus the local run start
                cnt[x] = 0;
                run += c;
            }
            if (t == 255) lstart[P] = run;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (key[j] != PC_KEY_SENTINEL) {
                unsigned int pos = lstart[q[j]] + atomicAdd(&cnt[q[j]], 1u);
                if (PACK) sorted[pos] = key[j] | ((unsigned long long)q[j] << PC_PACK_SHIFT);
                else { sorted[pos] = key[j]; qs[pos] = q[j]; }
            }
        }
        __syncthreads();
        unsigned int nv = lstart[P];
        for (unsigned int i = t; i < nv; i += 256) {
            unsigned long long v = sorted[i];
            unsigned int x = PACK ? (unsigned int)(v >> PC_PACK_SHIFT) : qs[i];
            kbuf[gbase[x] + i] = PACK ? (v & ((1ull << PC_PACK_SHIFT) - 1ull)) : v;
        }
        __syncthreads();
    }
}

// pass C: count one bucket after the other.  blockIdx = local_bucket * B + b: blocks are dispatched in index order,
// so at any time the resident blocks work on one or two adjacent buckets whose sub-tables (S slots) sit in L2.
// A bucket's k-mers arrive as n_seg segments (one per sending rank; a single segment on one GPU).
template <int ITEMS, bool CASFIRST, bool WATCH, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_count_part(const uint64_t* __restrict__ keys, const uint64_t* const* __restrict__ seg_base,
             const long long* __restrict__ seg_begin, const long long* __restrict__ seg_len,
             int n_seg, int pb, uint64_t S, unsigned int B, CountTable table, CountStats* __restrict__ stats,
             const Watch w_in)
{
    const Watch w = WATCH ? w_in : Watch{0u, nullptr, nullptr, 0ull};    // compile-time off: plain REDs, no extra registers
    unsigned int q = blockIdx.x / B, b = blockIdx.x % B;
    CountTable sub = table.sub((uint64_t)q * S);
    unsigned int new_keys = 0, max_probe = 0;
    const unsigned long long stride = (unsigned long long)B * 256u * ITEMS;
    for (int sg = 0; sg < n_seg; ++sg) {
        // seg_base != null: every segment lives in its sender's bucket buffer, mapped into this process over
        // NVLink (CUDA IPC).  The owner streams its peers' k-mers straight out of their HBM while it updates its
        // L2-resident sub-table: the exchange is fused into the count, no all-to-all, no receive buffer.
        const uint64_t* kbuf = (seg_base ? seg_base[(int64_t)q * n_seg + sg] : keys) + seg_begin[(int64_t)q * n_seg + sg];
        const unsigned long long end = (unsigned long long)seg_len[(int64_t)q * n_seg + sg];
        // software pipeline: the keys of the next iteration (a DRAM stream) are requested before this iteration's
        // table accesses (L2) are issued
        unsigned long long i0 = ((unsigned long long)b * 256u + threadIdx.x) * ITEMS;
        uint64_t nkey[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) nkey[j] = (i0 + j < end) ? __ldcs(kbuf + i0 + j) : PC_KEY_SENTINEL;
        for (; i0 < end; i0 += stride) {
            unsigned long long skey[ITEMS]; uint64_t slot[ITEMS]; unsigned long long cur[ITEMS];
            bool on[ITEMS];
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) {
                uint64_t key = nkey[j];
                on[j] = key != PC_KEY_SENTINEL;
                skey[j] = ~key;
                slot[j] = slot_in_part(hash64(key), pb, S);
            }
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) nkey[j] = (i0 + stride + j < end) ? __ldcs(kbuf + i0 + stride + j) : PC_KEY_SENTINEL;
            if (CASFIRST) {
#pragma unroll
                for (int j = 0; j < ITEMS; ++j) cur[j] = on[j] ? atomicCAS(sub.key(slot[j]), 0ull, skey[j]) : 0ull;
#pragma unroll
                for (int j = 0; j < ITEMS; ++j)
                    if (on[j]) count_insert_cas(sub, S, slot[j], skey[j], cur[j], new_keys, max_probe, stats, w);
            } else {
#pragma unroll
                for (int j = 0; j < ITEMS; ++j) cur[j] = on[j] ? ldcg_u64(sub.key(slot[j])) : 0ull;
                count_insert_batch<ITEMS>(sub, S, slot, skey, cur, on, new_keys, max_probe, stats, w);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        new_keys += __shfl_xor_sync(0xFFFFFFFFu, new_keys, o);
        max_probe = max(max_probe, __shfl_xor_sync(0xFFFFFFFFu, max_probe, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (new_keys) atomicAdd(&stats->distinct, (unsigned long long)new_keys);
        if (max_probe > 8) atomicMax(&stats->max_probe, (unsigned long long)max_probe);
    }
}

// merge (kmer, count) pairs received from other ranks into the owner's table (multi-GPU)
__global__ void __launch_bounds__(256)
k_merge(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts, int64_t n,
        CountTable table, uint64_t capacity, CountStats* __restrict__ stats)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = keys[i];
    unsigned int add = counts[i];
    unsigned long long skey = ~key;
    uint64_t slot = slot_of(hash64(key), capacity);
    unsigned int probes = 0;
    for (;;) {
        unsigned long long cur = ldcg_u64(table.key(slot));
        if (cur == 0ull) {
            cur = atomicCAS(table.key(slot), 0ull, skey);
            if (cur == 0ull) {                                // claimed: the slot already stands for one occurrence
                atomicAdd(&stats->distinct, 1ull);
                if (add > 1u) atomicAdd(table.cnt(slot), add - 1u);
                break;
            }
        }
        if (cur == skey) { atomicAdd(table.cnt(slot), add); break; }
        if (++probes > PC_PROBE_LIMIT || probes >= capacity) { atomicAdd(&stats->overflow, 1ull); break; }
        if (++slot == capacity) slot = 0;
    }
}

// ------------------------------------------------------------------------------------------------ K4 select
// Streaming scan of the table; survivors appended with one atomic per warp.  out_keys may be null (count only).
// Only the 4-byte counts are streamed; the key of a slot is read when its count passes the thresholds (an empty slot
// has count field 0 = "one occurrence", so for low >= 2 no key of an empty slot is ever touched).
__global__ void __launch_bounds__(256)
k_select(CountTable table, uint64_t capacity, unsigned int low, unsigned int high,
         uint64_t* __restrict__ out_keys, uint32_t* __restrict__ out_counts, unsigned long long out_cap,
         unsigned long long* __restrict__ n_out)
{
    // 4 slots per thread per iteration: one 128-bit streaming load of counts (the table arrays are 256-byte aligned)
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4;
    uint64_t cap_round = (capacity + 127) & ~127ull;
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < cap_round; i += stride) {
        unsigned int cnt[4] = {0u, 0u, 0u, 0u};
        if (i + 3 < capacity) {
            uint4 v = __ldcs(reinterpret_cast<const uint4*>(table.counts + i));
            cnt[0] = v.x; cnt[1] = v.y; cnt[2] = v.z; cnt[3] = v.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (i + j < capacity) cnt[j] = __ldcs(table.counts + i + j);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            bool keep = false; unsigned long long skey = 0;
            unsigned int c = cnt[j] + 1u;                                     // stored = occurrences - 1
            if (i + j < capacity && c >= low && c <= high) {
                skey = __ldcs(table.keys + i + j);
                keep = skey != 0ull;
            }
            unsigned int bal = __ballot_sync(0xFFFFFFFFu, keep);
            if (bal) {
                int lane = threadIdx.x & 31;
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(n_out, (unsigned long long)__popc(bal));
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (keep && out_keys) {
                    unsigned long long pos = base + __popc(bal & ((1u << lane) - 1u));
                    if (pos < out_cap) {
                        out_keys[pos] = ~skey;
                        if (out_counts) out_counts[pos] = c;
                    }
                }
            }
        }
    }
}

// owner histogram + scatter for the pre-aggregated multi-GPU exchange
__global__ void __launch_bounds__(256)
k_owner_hist(CountTable table, uint64_t capacity, unsigned int world, unsigned long long* __restrict__ hist /* world */)
{
    __shared__ unsigned int sh[64];
    if (threadIdx.x < 64) sh[threadIdx.x] = 0;
    __syncthreads();
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < capacity; i += stride) {
        unsigned long long skey = __ldcs(table.keys + i);
        if (skey != 0ull) atomicAdd(&sh[owner_of(~skey, world)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < world && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}
__global__ void __launch_bounds__(256)
k_owner_scatter(CountTable table, uint64_t capacity, unsigned int world,
                unsigned long long* __restrict__ cursor /* world, pre-set to bucket starts */,
                uint64_t* __restrict__ out_keys, uint32_t* __restrict__ out_counts)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < capacity; i += stride) {
        unsigned long long skey = __ldcs(table.keys + i);
        if (skey != 0ull) {
            unsigned int o = owner_of(~skey, world);
            unsigned int peers = __match_any_sync(__activemask(), o);
            int lane = threadIdx.x & 31;
            int leader = __ffs(peers) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(&cursor[o], (unsigned long long)__popc(peers));
            base = __shfl_sync(peers, base, leader);
            unsigned long long pos = base + __popc(peers & ((1u << lane) - 1u));
            out_keys[pos] = ~skey; out_counts[pos] = __ldcs(table.counts + i) + 1u;
        }
    }
}

// ------------------------------------------------------------------------------------------------ radix sort
// Stable LSD radix sort, 8-bit digits, warp-blocked ranges: warp gw owns the contiguous range
// [gw*ipw, (gw+1)*ipw).  hist is digit-major (hist[d*NW + gw]) so one exclusive scan over the flat array yields
// every (digit, warp) output base.  Ranking inside a warp uses match.any, no smem atomics in the scatter.
template <typename K>
__device__ __forceinline__ unsigned int digit_of(K key, int shift) { return (unsigned int)(key >> shift) & 0xFFu; }

template <typename K, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_rs_hist(const K* __restrict__ src, int64_t n, int64_t ipw, int shift, unsigned int* __restrict__ hist, int NW)
{
    __shared__ unsigned int h[WARPS][256];
    int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < WARPS * 256; i += WARPS * 32) (&h[0][0])[i] = 0;
    __syncthreads();
    int gw = blockIdx.x * WARPS + wid;
    int64_t beg = (int64_t)gw * ipw, end = min(n, beg + ipw);
    for (int64_t i = beg + lane; i < end; i += 32) atomicAdd(&h[wid][digit_of(src[i], shift)], 1u);
    __syncthreads();
    for (int d = lane; d < 256; d += 32) hist[(int64_t)d * NW + gw] = h[wid][d];
}

// single-CTA exclusive scan of `m` counters (m = 256*NW).  Warp q owns a contiguous segment and walks it with
// coalesced, 4x-unrolled loads; the 32 segment totals are scanned in shared memory.
template <typename T>
__global__ void __launch_bounds__(1024)
k_rs_scan(T* __restrict__ hist, int64_t m)                  // hist[m] receives the grand total
{
    __shared__ T wtot[32];
    int t = threadIdx.x, lane = t & 31, wq = t >> 5;
    int64_t seg = (((m + 31) / 32) + 127) & ~(int64_t)127;    // multiple of 128 = 4 tiles of 32
    int64_t beg = min(m, (int64_t)wq * seg), end = min(m, beg + seg);
    T s = 0;
    for (int64_t i = beg + lane; i < end; i += 128) {
        T a0 = hist[i];
        T a1 = (i + 32 < end) ? hist[i + 32] : (T)0;
        T a2 = (i + 64 < end) ? hist[i + 64] : (T)0;
        T a3 = (i + 96 < end) ? hist[i + 96] : (T)0;
        s += a0 + a1 + a2 + a3;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if (lane == 0) wtot[wq] = s;
    __syncthreads();
    if (wq == 0) {
        T v = wtot[lane], incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += u;
        }
        wtot[lane] = incl - v;
        if (lane == 31) hist[m] = incl;
    }
    __syncthreads();
    T carry = wtot[wq];
    for (int64_t i0 = beg; i0 < end; i0 += 128) {
        T v[4], incl[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { int64_t i = i0 + 32 * u + lane; v[u] = (i < end) ? hist[i] : (T)0; incl[u] = v[u]; }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                T x = __shfl_up_sync(0xFFFFFFFFu, incl[u], o);
                if (lane >= o) incl[u] += x;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int64_t i = i0 + 32 * u + lane;
            if (i < end) hist[i] = carry + incl[u] - v[u];
            carry += __shfl_sync(0xFFFFFFFFu, incl[u], 31);
        }
    }
}

template <typename K, int WARPS, bool HAS_VAL>
__global__ void __launch_bounds__(WARPS * 32)
k_rs_scatter(const K* __restrict__ src, K* __restrict__ dst, const uint32_t* __restrict__ vsrc,
             uint32_t* __restrict__ vdst, int64_t n, int64_t ipw, int shift,
             const unsigned int* __restrict__ hist, int NW)
{
    __shared__ unsigned int off[WARPS][256];
    int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int gw = blockIdx.x * WARPS + wid;
    for (int d = lane; d < 256; d += 32) off[wid][d] = hist[(int64_t)d * NW + gw];
    __syncwarp();
    int64_t beg = (int64_t)gw * ipw, end = min(n, beg + ipw);
    for (int64_t i0 = beg; i0 < end; i0 += 32) {
        int64_t i = i0 + lane;
        bool act = i < end;
        unsigned int amask = __ballot_sync(0xFFFFFFFFu, act);
        if (act) {
            K key = src[i];
            unsigned int d = digit_of(key, shift);
            unsigned int peers = __match_any_sync(amask, d);
            unsigned int rank = __popc(peers & ((1u << lane) - 1u));
            unsigned int base = off[wid][d];
            __syncwarp(amask);
            if (rank == 0) off[wid][d] = base + __popc(peers);
            __syncwarp(amask);
            dst[base + rank] = key;
            if (HAS_VAL) vdst[base + rank] = vsrc[i];
        }
    }
}

// drop consecutive duplicates of a sorted array (pc_set_selected); flags -> positions via one atomic per warp
// would break order, so this is a two-kernel exact compaction: per-CTA counts, host-free scan, scatter.
__global__ void __launch_bounds__(1024)
k_unique_sorted_single(const uint64_t* __restrict__ src, int64_t n, uint64_t* __restrict__ dst,
                       unsigned long long* __restrict__ n_out)
{
    // single CTA, ordered: used only for repeat sets (a few 1e6 keys at most per call)
    __shared__ unsigned int wsum[32];
    __shared__ unsigned long long carry;
    int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    if (t == 0) carry = 0;
    __syncthreads();
    for (int64_t i0 = 0; i0 < n; i0 += 1024) {
        int64_t i = i0 + t;
        bool flag = false; uint64_t v = 0;
        if (i < n) { v = src[i]; flag = (i == 0) || (src[i - 1] != v); }
        unsigned int bal = __ballot_sync(0xFFFFFFFFu, flag);
        if (lane == 0) wsum[wid] = __popc(bal);
        __syncthreads();
        unsigned int before = 0, total = 0;
        for (int x = 0; x < 32; ++x) { unsigned int c = wsum[x]; if (x < wid) before += c; total += c; }
        unsigned long long base = carry;
        if (flag) dst[base + before + __popc(bal & ((1u << lane) - 1u))] = v;
        __syncthreads();
        if (t == 0) carry = base + total;
        __syncthreads();
    }
    if (t == 0) *n_out = carry;
}

// number of adjacent duplicates in a sorted array (0 in the normal case: owners select disjoint key sets)
__global__ void __launch_bounds__(256)
k_count_dups(const uint64_t* __restrict__ sorted, int64_t n, unsigned long long* __restrict__ n_dups)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool dup = i > 0 && i < n && sorted[i] == sorted[i - 1];
    unsigned int bal = __ballot_sync(0xFFFFFFFFu, dup);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(n_dups, (unsigned long long)__popc(bal));
}

__global__ void __launch_bounds__(256)
k_canonicalize(uint64_t* __restrict__ keys, int64_t n, int k)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (keys[i] == PC_KEY_SENTINEL) return;                   // "not a k-mer" (pc_kmer_encode of a name with N): stays the sentinel
    uint64_t x = keys[i] & kmer_mask(k);
    uint64_t rc = revcomp_bits(x, k);
    keys[i] = x < rc ? x : rc;
}

// number of sentinel entries (the radix sort over 2k bits puts them last: no canonical-min k-mer has all 2k bits set)
__global__ void __launch_bounds__(256)
k_count_sentinel(const uint64_t* __restrict__ keys, int64_t n, unsigned long long* __restrict__ n_sent)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool s = i < n && keys[i] == PC_KEY_SENTINEL;
    unsigned int bal = __ballot_sync(0xFFFFFFFFu, s);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(n_sent, (unsigned long long)__popc(bal));
}

__global__ void __launch_bounds__(256)
k_stride_gather(const uint64_t* __restrict__ src, int64_t n_out, int64_t stride, uint64_t* __restrict__ dst)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_out) dst[i] = src[i * stride];
}

// ------------------------------------------------------------------------------------------------ K5 repeat table + probe
__global__ void __launch_bounds__(256)
k_rep_build(const uint64_t* __restrict__ sorted_keys, int64_t n, RepSlot* __restrict__ table, uint64_t capacity,
            unsigned short* __restrict__ fp /* one fingerprint per slot, zeroed = empty */)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = sorted_keys[i];
    unsigned long long skey = ~key;
    uint64_t slot = rep_slot(key, capacity);
    for (;;) {                                               // keys are unique and capacity >= 2n: terminates
        unsigned long long cur = atomicCAS(&table[slot].skey, 0ull, skey);
        if (cur == 0ull) { table[slot].col = (int)i; fp[slot] = fp_of(key); break; }
        if (++slot == capacity) slot = 0;
    }
}

// chunk of a packed word: chunk_word0 ascending.  Warp-uniform binary search for lane 0, then a short walk.
__device__ __forceinline__ int64_t chunk_of_word(const int64_t* __restrict__ chunk_word0, int64_t n_chunks,
                                                 int64_t w, int64_t w_lane0)
{
    int64_t c = find_le(chunk_word0, n_chunks, w_lane0);
    while (c + 1 < n_chunks && __ldg(chunk_word0 + c + 1) <= w) ++c;
    return c;
}

// Streams the packed genome once; every valid window probes the repeat table (one 32-byte sector, normally an
// L2 hit) and hits (column ids) are appended to the row's private region of `hits`.
template <int BATCH>
__global__ void __launch_bounds__(256)
k_probe(const uint64_t* __restrict__ words, const uint32_t* __restrict__ nmask, int64_t n_words, int k,
        const int64_t* __restrict__ chunk_word0, int64_t n_chunks, const int32_t* __restrict__ chunk_row,
        const int64_t* __restrict__ row_hit_base, const RepSlot* __restrict__ rtable, uint64_t rcap,
        const unsigned short* __restrict__ fp, int32_t* __restrict__ hits, unsigned int* __restrict__ row_cursor)
{
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int lane = threadIdx.x & 31;
    WordPair p = load_word_pair(words, nmask, w, n_words);
    int64_t w_lane0 = w - lane;
    if (w_lane0 >= n_words) return;                           // whole warp past the end
    int row = -1;
    if (w < n_words && p.m0 != 0xFFFFFFFFu) {
        int64_t c = chunk_of_word(chunk_word0, n_chunks, w, w_lane0);
        row = __ldg(chunk_row + c);
    }
    bool lane_on = row >= 0;
    unsigned int on_mask = __ballot_sync(0xFFFFFFFFu, lane_on);
    if (on_mask == 0u) return;
    int row0 = __shfl_sync(0xFFFFFFFFu, row, __ffs(on_mask) - 1);
    bool uniform = __all_sync(0xFFFFFFFFu, !lane_on || row == row0);

    Roller r;
    if (lane_on) r.init(p.w0, p.w1, p.m0, p.m1, k);
#pragma unroll 1
    for (int b0 = 0; b0 < 32; b0 += BATCH) {
        int cols[BATCH];
        unsigned int hmask = 0;                               // bit j: window b0+j hit repeat k-mer cols[j]
        if (lane_on) {
            unsigned long long skey[BATCH]; uint64_t slot[BATCH]; uint4 cur[BATCH];
            unsigned int vmask = 0;
#pragma unroll
            for (int j = 0; j < BATCH; ++j) {
                int b = b0 + j;
                uint64_t key = r.canonical();
                vmask |= (unsigned int)r.valid(b) << j;
                skey[j] = ~key;
                slot[j] = rep_slot(key, rcap);
                if (b < 31) r.step(b);
            }
            if (fp != nullptr) {
                // fingerprint-first probing (large repeat sets): the walk reads 2-byte fingerprints (16 slots per
                // sector, an array 8x smaller than the table, so it stays in L2 when the table does not) and touches a
                // 16-byte slot only when the fingerprint matches
                unsigned short f[BATCH];
#pragma unroll
                for (int j = 0; j < BATCH; ++j) f[j] = ((vmask >> j) & 1u) ? __ldg(fp + slot[j]) : (unsigned short)0;
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    cols[j] = 0;
                    if (!((vmask >> j) & 1u)) continue;
                    const unsigned short mine = fp_of(~skey[j]);
                    unsigned short ff = f[j];
                    uint64_t s = slot[j];
                    while (ff != 0) {
                        if (ff == mine) {
                            uint4 c2 = __ldg(reinterpret_cast<const uint4*>(rtable + s));
                            unsigned long long ck = ((unsigned long long)c2.y << 32) | c2.x;
                            if (ck == skey[j]) { cols[j] = (int)c2.z; hmask |= 1u << j; break; }
                        }
                        if (++s == rcap) s = 0;
                        ff = __ldg(fp + s);
                    }
                }
                vmask = 0;                                    // done: skip the direct walk below
            }
#pragma unroll
            for (int j = 0; j < BATCH; ++j)
                cur[j] = ((vmask >> j) & 1u) ? __ldg(reinterpret_cast<const uint4*>(rtable + slot[j])) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int j = 0; j < BATCH; ++j) {
                if (fp == nullptr) cols[j] = 0;
                if (!((vmask >> j) & 1u)) continue;
                unsigned long long ck = ((unsigned long long)cur[j].y << 32) | cur[j].x;
                int col = (int)cur[j].z;
                uint64_t s = slot[j];
                while (ck != 0ull && ck != skey[j]) {          // linear probing until empty or found
                    if (++s == rcap) s = 0;
                    uint4 c2 = __ldg(reinterpret_cast<const uint4*>(rtable + s));
                    ck = ((unsigned long long)c2.y << 32) | c2.x; col = (int)c2.z;
                }
                if (ck != 0ull) { cols[j] = col; hmask |= 1u << j; }
            }
        }
        int nh = __popc(hmask);
        // append this batch's hits: one atomic per warp when the warp sits in one row (the normal case)
        unsigned int any = __ballot_sync(0xFFFFFFFFu, nh > 0);
        if (any == 0u) continue;
        int32_t* dst = nullptr;
        if (uniform) {
            int incl = nh;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += v;
            }
            int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(&row_cursor[row0], (unsigned int)total);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (nh > 0) dst = hits + __ldg(row_hit_base + row0) + base + (incl - nh);
        } else if (nh > 0) {
            unsigned int base = atomicAdd(&row_cursor[row], (unsigned int)nh);
            dst = hits + __ldg(row_hit_base + row) + base;
        }
        if (nh > 0) {
#pragma unroll
            for (int j = 0; j < BATCH; ++j)
                if ((hmask >> j) & 1u) dst[__popc(hmask & ((1u << j) - 1u))] = cols[j];
        }
    }
}

// Stream variant of the probe: when the partitioned count already materialised the canonical k-mer of every window
// (k_extract), K5 re-reads that stream (8 coalesced bytes per window) instead of re-running the roller.
template <int BATCH>
__global__ void __launch_bounds__(256)
k_probe_stream(const uint64_t* __restrict__ stream, int64_t n_words,
               const int64_t* __restrict__ chunk_word0, int64_t n_chunks, const int32_t* __restrict__ chunk_row,
               const int64_t* __restrict__ row_hit_base, const RepSlot* __restrict__ rtable, uint64_t rcap,
               const unsigned short* __restrict__ fp, const unsigned int* __restrict__ gbits, unsigned int gnbits,
               int32_t* __restrict__ hits, unsigned int* __restrict__ row_cursor)
{
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int lane = threadIdx.x & 31;
    int64_t w_lane0 = w - lane;
    if (w_lane0 >= n_words) return;                           // whole warp past the end
    int row = -1;
    if (w < n_words) {
        int64_t c = chunk_of_word(chunk_word0, n_chunks, w, w_lane0);
        row = __ldg(chunk_row + c);
    }
    bool lane_on = row >= 0;
    unsigned int on_mask = __ballot_sync(0xFFFFFFFFu, lane_on);
    if (on_mask == 0u) return;
    int row0 = __shfl_sync(0xFFFFFFFFu, row, __ffs(on_mask) - 1);
    bool uniform = __all_sync(0xFFFFFFFFu, !lane_on || row == row0);
    const uint64_t* src = stream + w_lane0 * 32 + lane;
    // software pipeline: the stream k-mers of the next batch are requested (DRAM) before this batch's table probes (L2)
    // are issued -- ncu showed a third of the stall samples on the first use of the stream loads
    uint64_t nkey[BATCH];
#pragma unroll
    for (int j = 0; j < BATCH; ++j) nkey[j] = lane_on ? __ldcs(src + j * 32) : PC_KEY_SENTINEL;
#pragma unroll 1
    for (int b0 = 0; b0 < 32; b0 += BATCH) {
        int cols[BATCH];
        unsigned int hmask = 0;
        uint64_t kcur[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j) kcur[j] = nkey[j];
        if (lane_on && b0 + BATCH < 32) {
#pragma unroll
            for (int j = 0; j < BATCH; ++j) nkey[j] = __ldcs(src + (b0 + BATCH + j) * 32);
        }
        if (lane_on) {
            unsigned long long skey[BATCH]; uint64_t slot[BATCH]; uint4 cur[BATCH];
            unsigned int vmask = 0;
#pragma unroll
            for (int j = 0; j < BATCH; ++j) {
                uint64_t key = kcur[j];
                vmask |= (unsigned int)(key != PC_KEY_SENTINEL) << j;
                skey[j] = ~key;
                slot[j] = rep_slot(key, rcap);
            }
            if (gbits != nullptr) {
                // repeat sets too large for the shared-memory filter have a repeat table beyond L2: one bit per hash
                // bucket in a bitmap that DOES stay in L2 (16 bits per repeat k-mer) keeps ~2/3 of the windows from
                // touching DRAM at all
                unsigned int wbits[BATCH], fb[BATCH];
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    fb[j] = gbit_of(~skey[j], gnbits);
                    wbits[j] = ((vmask >> j) & 1u) ? __ldg(gbits + (fb[j] >> 5)) : 0u;
                }
#pragma unroll
                for (int j = 0; j < BATCH; ++j)
                    if (!((wbits[j] >> (fb[j] & 31)) & 1u)) vmask &= ~(1u << j);
            }
            if (fp != nullptr) {
                // fingerprint-first probing (large repeat sets): the walk reads 2-byte fingerprints (16 slots per
                // sector, an array 8x smaller than the table, so it stays in L2 when the table does not) and touches a
                // 16-byte slot only when the fingerprint matches
                unsigned short f[BATCH];
#pragma unroll
                for (int j = 0; j < BATCH; ++j) f[j] = ((vmask >> j) & 1u) ? __ldg(fp + slot[j]) : (unsigned short)0;
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    cols[j] = 0;
                    if (!((vmask >> j) & 1u)) continue;
                    const unsigned short mine = fp_of(~skey[j]);
                    unsigned short ff = f[j];
                    uint64_t s = slot[j];
                    while (ff != 0) {
                        if (ff == mine) {
                            uint4 c2 = __ldg(reinterpret_cast<const uint4*>(rtable + s));
                            unsigned long long ck = ((unsigned long long)c2.y << 32) | c2.x;
                            if (ck == skey[j]) { cols[j] = (int)c2.z; hmask |= 1u << j; break; }
                        }
                        if (++s == rcap) s = 0;
                        ff = __ldg(fp + s);
                    }
                }
                vmask = 0;                                    // done: skip the direct walk below
            }
#pragma unroll
            for (int j = 0; j < BATCH; ++j)
                cur[j] = ((vmask >> j) & 1u) ? __ldg(reinterpret_cast<const uint4*>(rtable + slot[j])) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int j = 0; j < BATCH; ++j) {
                if (fp == nullptr) cols[j] = 0;
                if (!((vmask >> j) & 1u)) continue;
                unsigned long long ck = ((unsigned long long)cur[j].y << 32) | cur[j].x;
                int col = (int)cur[j].z;
                uint64_t s = slot[j];
                while (ck != 0ull && ck != skey[j]) {
                    if (++s == rcap) s = 0;
                    uint4 c2 = __ldg(reinterpret_cast<const uint4*>(rtable + s));
                    ck = ((unsigned long long)c2.y << 32) | c2.x; col = (int)c2.z;
                }
                if (ck != 0ull) { cols[j] = col; hmask |= 1u << j; }
            }
        }
        int nh = __popc(hmask);
        unsigned int any = __ballot_sync(0xFFFFFFFFu, nh > 0);
        if (any == 0u) continue;
        int32_t* dst = nullptr;
        if (uniform) {
            int incl = nh;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += v;
            }
            int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(&row_cursor[row0], (unsigned int)total);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (nh > 0) dst = hits + __ldg(row_hit_base + row0) + base + (incl - nh);
        } else if (nh > 0) {
            unsigned int base = atomicAdd(&row_cursor[row], (unsigned int)nh);
            dst = hits + __ldg(row_hit_base + row) + base;
        }
        if (nh > 0) {
#pragma unroll
            for (int j = 0; j < BATCH; ++j)
                if ((hmask >> j) & 1u) dst[__popc(hmask & ((1u << j) - 1u))] = cols[j];
        }
    }
}

// ------------------------------------------------------------------------------------------------ K6 rows -> CSR
// One CTA per matrix row.  The row's hits (column ids, unordered, with multiplicity) are sorted with the same
// stable LSD scheme, CTA-local: warp wq owns a contiguous slice, hist[wq][256] lives in shared memory, data
// ping-pongs between the row's slices of two global buffers (L2 resident: a 75 kb chunk is <= 300 KB).
template <int WARPS, int DB>
__global__ void __launch_bounds__(WARPS * 32)
k_row_sort(int32_t* buf_a, int32_t* buf_b, const int64_t* __restrict__ row_hit_base,
           const unsigned int* __restrict__ row_cursor, int n_passes, int64_t* __restrict__ row_nnz)
{
    // DB digit bits per pass (8, 10 or 11): 20-bit column ids sort in two passes of 10 bits instead of three of 8
    constexpr int BINS = 1 << DB;
    extern __shared__ unsigned int rs_smem[];
    unsigned int (*h)[BINS] = reinterpret_cast<unsigned int (*)[BINS]>(rs_smem);   // [WARPS][BINS]
    unsigned int* tot = rs_smem + WARPS * BINS;                                    // [BINS]
    __shared__ unsigned int red[WARPS];
    __shared__ unsigned int carry_s[BINS / 32 + 1];
    int row = blockIdx.x;
    int t = threadIdx.x, wid = t >> 5, lane = t & 31;
    unsigned int n = row_cursor[row];
    int64_t base = row_hit_base[row];
    int32_t* src = buf_a + base;
    int32_t* dst = buf_b + base;
    unsigned int ipw = ((n + WARPS - 1) / WARPS + 31u) & ~31u;
    unsigned int beg = min(n, wid * ipw), end = min(n, beg + ipw);
    for (int pass = 0; pass < n_passes; ++pass) {
        int shift = DB * pass;
        for (int i = t; i < WARPS * BINS; i += WARPS * 32) rs_smem[i] = 0;
        __syncthreads();
        for (unsigned int i = beg + lane; i < end; i += 32)
            atomicAdd(&h[wid][((unsigned int)src[i] >> shift) & (BINS - 1)], 1u);
        __syncthreads();
        for (int d = t; d < BINS; d += WARPS * 32) {          // per digit: exclusive scan over warps
            unsigned int s = 0;
#pragma unroll
            for (int q = 0; q < WARPS; ++q) { unsigned int v = h[q][d]; h[q][d] = s; s += v; }
            tot[d] = s;
        }
        __syncthreads();
        // exclusive scan of the BINS digit totals: each warp scans 32-wide groups, then the group sums are chained
        for (int g = wid; g < BINS / 32; g += WARPS) {
            unsigned int v = tot[g * 32 + lane], incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned int u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += u;
            }
            tot[g * 32 + lane] = incl - v;
            if (lane == 31) carry_s[g] = incl;
        }
        __syncthreads();
        if (wid == 0) {
            unsigned int run = 0;
            for (int g0 = 0; g0 < BINS / 32; g0 += 32) {
                int g = g0 + lane;
                unsigned int v = (g < BINS / 32) ? carry_s[g] : 0u, incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    unsigned int u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (lane >= o) incl += u;
                }
                if (g < BINS / 32) carry_s[g] = run + incl - v;
                run += __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
        }
        __syncthreads();
        for (int d = lane; d < BINS; d += 32) h[wid][d] += tot[d] + carry_s[d >> 5];
        __syncwarp();
        for (unsigned int i0 = beg; i0 < end; i0 += 32) {
            unsigned int i = i0 + lane;
            bool act = i < end;
            unsigned int amask = __ballot_sync(0xFFFFFFFFu, act);
            if (act) {
                int32_t key = src[i];
                unsigned int d = ((unsigned int)key >> shift) & (BINS - 1);
                unsigned int peers = __match_any_sync(amask, d);
                unsigned int rank = __popc(peers & ((1u << lane) - 1u));
                unsigned int ob = h[wid][d];
                __syncwarp(amask);
                if (rank == 0) h[wid][d] = ob + __popc(peers);
                __syncwarp(amask);
                dst[ob + rank] = key;
            }
        }
        __syncthreads();                                      // dst complete and visible to the whole CTA
        int32_t* tmp = src; src = dst; dst = tmp;
    }
    // number of distinct columns in the sorted row
    unsigned int cnt = 0;
    for (unsigned int i = t; i < n; i += WARPS * 32) cnt += (i == 0 || src[i] != src[i - 1]) ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    if (lane == 0) red[wid] = cnt;
    __syncthreads();
    if (t == 0) {
        unsigned int s = 0;
        for (int q = 0; q < WARPS; ++q) s += red[q];
        row_nnz[row] = (int64_t)s;
    }
}

// exclusive scan of row_nnz -> indptr (R+1).  Single CTA; R is at most a few 1e5.
__global__ void __launch_bounds__(1024)
k_indptr(const int64_t* __restrict__ row_nnz, int64_t n_rows, int64_t* __restrict__ indptr)
{
    __shared__ unsigned long long part[1024];
    int t = threadIdx.x;
    int64_t per = (n_rows + 1023) / 1024;
    int64_t beg = (int64_t)t * per, end = min(n_rows, beg + per);
    unsigned long long s = 0;
    for (int64_t i = beg; i < end; ++i) s += (unsigned long long)row_nnz[i];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        unsigned long long v = (t >= o) ? part[t - o] : 0ull;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    unsigned long long run = part[t] - s;
    for (int64_t i = beg; i < end; ++i) { indptr[i] = (int64_t)run; run += (unsigned long long)row_nnz[i]; }
    if (t == 1023) indptr[n_rows] = (int64_t)part[1023];
}

// run-length encode each sorted row into CSR (indices ascending, data = multiplicity) and accumulate the
// column document frequency df (= stored entries per column, what TfidfTransformer calls df).
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_row_emit(const int32_t* __restrict__ sorted, const int64_t* __restrict__ row_hit_base,
           const unsigned int* __restrict__ row_cursor, const int64_t* __restrict__ indptr,
           int32_t* __restrict__ indices, float* __restrict__ data, int* __restrict__ df)
{
    __shared__ unsigned int wsum[THREADS / 32];
    __shared__ unsigned int carry_s;
    int row = blockIdx.x;
    int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    unsigned int n = row_cursor[row];
    const int32_t* a = sorted + row_hit_base[row];
    int64_t out0 = indptr[row];
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (unsigned int i0 = 0; i0 < n; i0 += THREADS) {
        unsigned int i = i0 + t;
        bool flag = false; int32_t v = 0;
        if (i < n) { v = a[i]; flag = (i == 0) || (a[i - 1] != v); }
        unsigned int bal = __ballot_sync(0xFFFFFFFFu, flag);
        if (lane == 0) wsum[wid] = __popc(bal);
        __syncthreads();
        unsigned int before = 0, total = 0;
#pragma unroll
        for (int x = 0; x < THREADS / 32; ++x) { unsigned int c = wsum[x]; if (x < wid) before += c; total += c; }
        unsigned int base = carry_s;
        if (flag) {
            // run end: short linear look-ahead, then binary search (runs are usually 1-3 long)
            unsigned int e = i + 1;
            int steps = 0;
            while (e < n && steps < 4 && a[e] == v) { ++e; ++steps; }
            if (e < n && a[e] == v) {
                unsigned int lo = e, hi = n;                  // a[lo] == v, find first index with a[] != v
                while (hi - lo > 1) { unsigned int mid = lo + ((hi - lo) >> 1); if (a[mid] == v) lo = mid; else hi = mid; }
                e = hi;
            }
            int64_t pos = out0 + base + before + __popc(bal & ((1u << lane) - 1u));
            indices[pos] = v;
            data[pos] = (float)(e - i);
            atomicAdd(&df[v], 1);
        }
        __syncthreads();
        if (t == 0) carry_s = base + total;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ K7 tf-idf
// idf_j = ln(float32(n+1) / float32(df_j+1)) + 1, all float32 like sklearn with float32 input
// (text.py: df.astype(dtype); full_like/df; np.log; += 1).
__global__ void __launch_bounds__(256)
k_idf(const int* __restrict__ df, int64_t n_cols, float n_plus_1, float* __restrict__ idf)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_cols) return;
    float q = __fdiv_rn(n_plus_1, (float)df[j] + 1.0f);
    idf[j] = (float)log((double)q) + 1.0f;                   // correctly rounded float32 log, then float32 add
}

// One CTA per row: x = count*idf (float32), sum of float32 squares accumulated in double, x /= sqrt(sum) in
// double and rounded once (sparsefuncs_fast.pyx:_inplace_csr_row_normalize_l2).  Zero rows stay zero.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_tfidf(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ data,
        const float* __restrict__ idf, float* __restrict__ out, double* __restrict__ row_norm)
{
    __shared__ double red[THREADS / 32];
    __shared__ double norm_s;
    int row = blockIdx.x;
    int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    int64_t beg = indptr[row], end = indptr[row + 1];
    double acc = 0.0;
    for (int64_t i = beg + t; i < end; i += THREADS) {
        float x = __fmul_rn(data[i], __ldg(idf + indices[i]));
        out[i] = x;
        acc += (double)__fmul_rn(x, x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if (lane == 0) red[wid] = acc;
    __syncthreads();
    if (t == 0) {
        double s = 0.0;
        for (int q = 0; q < THREADS / 32; ++q) s += red[q];
        norm_s = sqrt(s);
        if (row_norm) row_norm[row] = norm_s;                 // with idf, all a host needs to rebuild the values (pc_tfidf_factors_get)
    }
    __syncthreads();
    double nrm = norm_s;
    if (nrm == 0.0) return;
    for (int64_t i = beg + t; i < end; i += THREADS) out[i] = (float)((double)out[i] / nrm);
}

// Compact hand-off of the matrix over PCIe: the counts are small integers (at most the windows of a chunk), so they travel
// as uint16 instead of float32; the rare entry >= 65535 (a poly-A chunk) is stored as 65535 and listed apart.
__global__ void __launch_bounds__(256)
k_data_u16(const float* __restrict__ data, int64_t n, unsigned short* __restrict__ out, unsigned long long* __restrict__ n_esc,
           int64_t* __restrict__ esc_pos, float* __restrict__ esc_val, unsigned long long esc_cap)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = data[i];
        if (v >= 65535.0f) {
            const unsigned long long p = atomicAdd(n_esc, 1ull);
            if (p < esc_cap) { esc_pos[p] = i; esc_val[p] = v; }
            out[i] = 65535;
        } else out[i] = (unsigned short)v;
    }
}

// Densest hand-off: one byte per column id (the gap to the previous column of the row; the first entry of a row and any gap
// >= 255 are listed apart with their absolute column) and one byte per count (>= 255 listed apart).  Column ids of a row
// ascend with a mean gap of n_selected / (entries per row) ~ 50, counts are mostly 1-3: 2 bytes per entry instead of 12.
// GT = unsigned char (escape 255) or unsigned short (escape 65535): with a large column universe the mean gap of a row
// passes 255 (8 ranks, 5.6 M columns, 21 k entries per row: mean gap 257) and half of the one-byte gaps would escape at 12
// bytes each -- measured 0.99 GB per rank and step instead of 0.29; two-byte gaps then cost 3 bytes per entry.
template <int THREADS, typename GT>
__global__ void __launch_bounds__(THREADS)
k_csr_pack8(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ data,
            GT* __restrict__ dcol, unsigned char* __restrict__ cnt8,
            unsigned long long* __restrict__ n_esc /* [0] columns, [1] counts */, int64_t* __restrict__ col_pos,
            int32_t* __restrict__ col_val, unsigned long long col_cap, int64_t* __restrict__ cnt_pos,
            float* __restrict__ cnt_val, unsigned long long cnt_cap)
{
    const int64_t row = blockIdx.x;
    const int64_t beg = indptr[row], end = indptr[row + 1];
    for (int64_t i = beg + threadIdx.x; i < end; i += THREADS) {
        const int32_t c = indices[i];
        constexpr int64_t ESC = sizeof(GT) == 1 ? 255 : 65535;
        const int64_t gap = i == beg ? ESC : (int64_t)c - (int64_t)indices[i - 1];
        if (gap >= ESC) {
            const unsigned long long p = atomicAdd(&n_esc[0], 1ull);
            if (p < col_cap) { col_pos[p] = i; col_val[p] = c; }
            dcol[i] = (GT)ESC;
        } else dcol[i] = (GT)gap;
        const float v = data[i];
        if (v >= 255.0f) {
            const unsigned long long p = atomicAdd(&n_esc[1], 1ull);
            if (p < cnt_cap) { cnt_pos[p] = i; cnt_val[p] = v; }
            cnt8[i] = 255;
        } else cnt8[i] = (unsigned char)v;
    }
}

// Two bytes per entry for a large column universe: (gap << 4) | count, gap < 4095 and count < 15, else the field holds
// its maximum and the value is listed apart (column escapes: absolute column; count escapes: the count).
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_csr_pack_g12c4(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ data,
                 unsigned short* __restrict__ packed, unsigned long long* __restrict__ n_esc /* [0] columns, [1] counts */,
                 int64_t* __restrict__ col_pos, int32_t* __restrict__ col_val, unsigned long long col_cap,
                 int64_t* __restrict__ cnt_pos, float* __restrict__ cnt_val, unsigned long long cnt_cap)
{
    const int64_t row = blockIdx.x;
    const int64_t beg = indptr[row], end = indptr[row + 1];
    for (int64_t i = beg + threadIdx.x; i < end; i += THREADS) {
        const int32_t c = indices[i];
        int64_t gap = i == beg ? 4095 : (int64_t)c - (int64_t)indices[i - 1];
        if (gap >= 4095) {
            const unsigned long long p = atomicAdd(&n_esc[0], 1ull);
            if (p < col_cap) { col_pos[p] = i; col_val[p] = c; }
            gap = 4095;
        }
        const float v = data[i];
        unsigned int cn = (unsigned int)v;
        if (v >= 15.0f) {
            const unsigned long long p = atomicAdd(&n_esc[1], 1ull);
            if (p < cnt_cap) { cnt_pos[p] = i; cnt_val[p] = v; }
            cn = 15u;
        }
        packed[i] = (unsigned short)(((unsigned int)gap << 4) | cn);
    }
}

// ------------------------------------------------------------------------------------------------ N1 subgenome extraction
// Row N1 of the scope table (polycracker.py:692-865): per-subgenome k-mer dictionaries, the differential-k-mer ratio
// test, and the map-back of the differential k-mers onto every chunk.  All dictionaries are RepSlot tables (key -> int).
constexpr int kMaxGroups = 8;                                  // subgenomes per session (one mask byte per k-mer)
constexpr int kMaxSubK = 8;                                    // k-mer lengths per session (reference default: 23,31)

struct GroupTables { const RepSlot* t[kMaxGroups]; uint64_t cap[kMaxGroups]; int n; };

// value stored for `key`, 0 when absent (dictionary values are counts >= 1 or non-empty masks)
__device__ __forceinline__ int kv_find(const RepSlot* __restrict__ table, uint64_t cap, uint64_t key)
{
    if (cap == 0) return 0;
    const unsigned long long skey = ~key;
    uint64_t s = rep_slot(key, cap);
    for (;;) {
        uint4 c = __ldg(reinterpret_cast<const uint4*>(table + s));
        unsigned long long ck = ((unsigned long long)c.y << 32) | c.x;
        if (ck == 0ull) return 0;
        if (ck == skey) return (int)c.z;
        if (++s == cap) s = 0;
    }
}

// kmercounttodict (polycracker.py:722-733): unique keys, value = count
__global__ void __launch_bounds__(256)
k_kv_build(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t n,
           RepSlot* __restrict__ table, uint64_t cap)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = keys[i];
    unsigned long long skey = ~key;
    uint64_t s = rep_slot(key, cap);
    for (;;) {
        unsigned long long cur = atomicCAS(&table[s].skey, 0ull, skey);
        if (cur == 0ull) { table[s].col = (int)vals[i]; break; }
        if (++s == cap) s = 0;
    }
}

// compareKmers (polycracker.py:735-783): a k-mer of this subgenome is differential when its count divided by the count
// in ANY other subgenome exceeds the threshold.  The reference is Python 2: int / int floors when the other dictionary
// holds the k-mer (pc.py:765, 778), and is a float division by default_kmercount_value (a float) when it does not.
__global__ void __launch_bounds__(256)
k_diff_test(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts, int64_t n, GroupTables others,
            double ratio_threshold, double default_value, uint64_t* __restrict__ out, unsigned long long* counter)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = keys[i];
    uint32_t v1 = counts[i];
    bool pass = false;
    for (int o = 0; o < others.n; ++o) {
        int v2 = kv_find(others.t[o], others.cap[o], key);
        if (v2 > 0) pass |= (double)(v1 / (uint32_t)v2) > ratio_threshold;
        else pass |= ((double)v1 / default_value) > ratio_threshold;
    }
    if (pass) out[atomicAdd(counter, 1ull)] = key;
}

// header values of `>kmer.val1.others` (pc.py:767, 781): the count of each differential k-mer in every subgenome
__global__ void __launch_bounds__(256)
k_diff_counts(const uint64_t* __restrict__ keys, int64_t n, GroupTables all, uint32_t* __restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = keys[i];
    for (int g = 0; g < all.n; ++g) out[i * all.n + g] = (uint32_t)kv_find(all.t[g], all.cap[g], key);
}

// union of the differential sets: key -> bit mask of the subgenomes it is differential for
__global__ void __launch_bounds__(256)
k_mask_build(const uint64_t* __restrict__ keys, int64_t n, int bit, RepSlot* __restrict__ table, uint64_t cap)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = keys[i];
    unsigned long long skey = ~key;
    uint64_t s = rep_slot(key, cap);
    for (;;) {
        unsigned long long cur = atomicCAS(&table[s].skey, 0ull, skey);
        if (cur == 0ull || cur == skey) { atomicOr(&table[s].col, bit); break; }
        if (++s == cap) s = 0;
    }
}

// N4, k-mer x genome count matrix (`bio_hyp_class`, utilities.py:177-284): over the UNION of the genomes' dumped k-mers,
// keep a k-mer when max / clip(min, default) >= ratio (Python-2 integer arrays: the division floors; min is 0 when any
// genome lacks the k-mer, because the matrix is sparse with implicit zeros) and max >= max_val (utilities.py:255-258).
// Run once per genome g over its dictionary; a k-mer is handled by the FIRST genome that holds it, so every union member
// is tested exactly once.
__global__ void __launch_bounds__(256)
k_union_test(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts, int64_t n, int g, GroupTables all,
             long long ratio_threshold, long long default_value, long long max_val,
             uint64_t* __restrict__ out, unsigned long long* __restrict__ n_out, unsigned long long* __restrict__ n_union)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool first = false, keep = false;
    uint64_t key = 0;
    if (i < n) {
        key = keys[i];
        first = true;
        long long mx = counts[i], mn = counts[i];
        for (int o = 0; o < all.n; ++o) {
            if (o == g) continue;
            long long v = kv_find(all.t[o], all.cap[o], key);
            if (o < g && v > 0) { first = false; break; }           // an earlier genome owns this k-mer
            mx = max(mx, v); mn = min(mn, v);
        }
        if (first) {
            long long clipped = max(mn, default_value);
            keep = clipped > 0 && (mx / clipped) >= ratio_threshold && mx >= max_val;
        }
    }
    unsigned int bal_u = __ballot_sync(0xFFFFFFFFu, first);
    unsigned int bal_k = __ballot_sync(0xFFFFFFFFu, keep);
    int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) {
        if (bal_u) atomicAdd(n_union, (unsigned long long)__popc(bal_u));
        if (bal_k) base = atomicAdd(n_out, (unsigned long long)__popc(bal_k));
    }
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (keep) out[base + __popc(bal_k & ((1u << lane) - 1u))] = key;
}

struct SubKSet { const RepSlot* t[kMaxSubK]; uint64_t cap[kMaxSubK]; int k[kMaxSubK]; int n; };

// writeBlast + blast2bed3 (polycracker.py:786-865): bbmap perfectmode reports every exact occurrence (either strand) of
// every differential k-mer; blast2bed3 turns each into the 1-base interval at its leftmost position and `bedtools
// coverage` reports, per chunk, the number of bases covered at least once (column 6 of the 4-column windows) -- i.e.
// the number of DISTINCT window starts where a differential k-mer of that subgenome (of any of the lengths) matches.
// One thread owns the 32 window starts of one packed word for every k, so the union over lengths is a register OR.
template <int BATCH>
__global__ void __launch_bounds__(256)
k_probe_groups(const uint64_t* __restrict__ words, const uint32_t* __restrict__ nmask, int64_t n_words, SubKSet ks,
               const int64_t* __restrict__ chunk_word0, int64_t n_chunks, int n_groups, int* __restrict__ coverage)
{
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int lane = threadIdx.x & 31;
    WordPair p = load_word_pair(words, nmask, w, n_words);
    int64_t w_lane0 = w - lane;
    if (w_lane0 >= n_words) return;
    bool lane_on = w < n_words && p.m0 != 0xFFFFFFFFu;
    int64_t chunk = -1;
    if (lane_on) chunk = chunk_of_word(chunk_word0, n_chunks, w, w_lane0);
    unsigned int on_mask = __ballot_sync(0xFFFFFFFFu, lane_on);
    if (on_mask == 0u) return;
    int64_t chunk0 = __shfl_sync(0xFFFFFFFFu, chunk, __ffs(on_mask) - 1);
    bool uniform = __all_sync(0xFFFFFFFFu, !lane_on || chunk == chunk0);

    unsigned int hit[kMaxGroups];                             // bit b of hit[g]: window start b matched subgenome g
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) hit[g] = 0u;
    if (lane_on) {
#pragma unroll 1
        for (int q = 0; q < ks.n; ++q) {
            const RepSlot* __restrict__ table = ks.t[q];
            const uint64_t cap = ks.cap[q];
            if (cap == 0) continue;
            Roller r;
            r.init(p.w0, p.w1, p.m0, p.m1, ks.k[q]);
#pragma unroll 1
            for (int b0 = 0; b0 < 32; b0 += BATCH) {
                unsigned long long skey[BATCH]; uint64_t slot[BATCH]; uint4 cur[BATCH];
                unsigned int vmask = 0;
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    int b = b0 + j;
                    uint64_t key = r.canonical();
                    vmask |= (unsigned int)r.valid(b) << j;
                    skey[j] = ~key;
                    slot[j] = rep_slot(key, cap);
                    if (b < 31) r.step(b);
                }
#pragma unroll
                for (int j = 0; j < BATCH; ++j)
                    cur[j] = ((vmask >> j) & 1u) ? __ldg(reinterpret_cast<const uint4*>(table + slot[j])) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    if (!((vmask >> j) & 1u)) continue;
                    unsigned long long ck = ((unsigned long long)cur[j].y << 32) | cur[j].x;
                    unsigned int m = cur[j].z;
                    uint64_t s = slot[j];
                    while (ck != 0ull && ck != skey[j]) {
                        if (++s == cap) s = 0;
                        uint4 c2 = __ldg(reinterpret_cast<const uint4*>(table + s));
                        ck = ((unsigned long long)c2.y << 32) | c2.x; m = c2.z;
                    }
                    if (ck != 0ull) {
#pragma unroll
                        for (int g = 0; g < kMaxGroups; ++g) hit[g] |= ((m >> g) & 1u) << (b0 + j);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) {
        if (g >= n_groups) break;
        int cnt = __popc(hit[g]);
        if (uniform) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
            if (lane == 0 && cnt) atomicAdd(&coverage[chunk0 * n_groups + g], cnt);
        } else if (cnt) {
            atomicAdd(&coverage[chunk * n_groups + g], cnt);
        }
    }
}

}  // namespace pc
