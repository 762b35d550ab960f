// K3, shared-memory path: two-level radix partition, then every sub-bucket is counted in a shared-memory hash table.
//
// The L2-resident bucket count (k_count_part) is bound by random L2 sector operations (~4 per insert, half of them
// across the die-to-die fabric; profiles/r01_kbench_count.md).  This path trades them for one more streaming pass:
// each level-1 bucket (top pb bits of the hash) is split again into NB2 sub-buckets of a few thousand k-mers, small
// enough that ALL distinct k-mers of a sub-bucket fit a shared-memory table.  One CTA then streams a sub-bucket
// (contiguous, coalesced), inserts with shared-memory atomics, and writes out only the k-mers whose count reaches
// keep_min as a compact (k-mer, count) list -- no global table, no table memset, and the threshold scan (K4) reads the
// list instead of 12 B x capacity.  A sub-bucket whose distinct k-mers do not fit (adversarial input, capped NB2) is
// reported in an overflow list and counted by the global-table kernel instead, so the result never depends on sizing.
#pragma once
#include "kernels.cuh"

namespace pc {

#define PC_SUB_MAX 2048                                   // sub-buckets per level-1 bucket (smem arrays of the scatter)

// sub-bucket of a k-mer inside its level-1 bucket: the 32 hash bits below the bucket bits, scaled to [0, NB2)
__device__ __forceinline__ unsigned int sub_of(uint64_t h, int pb, unsigned int NB2) {
    return (unsigned int)((((h << pb) >> 32) * (uint64_t)NB2) >> 32);
}

// A level-1 bucket q arrives as n_seg segments (one per sending rank; one on a single GPU), pair p = q * n_seg + sg.
// Tiles never span two pairs, so every key of a tile belongs to level-1 bucket q: tile_start[p] = first tile of pair p.
struct SegTile { const uint64_t* src; int64_t lim; int64_t q; };
template <int TILE>
__device__ __forceinline__ SegTile seg_tile(const uint64_t* __restrict__ keys, const uint64_t* const* __restrict__ seg_base,
                                            const int64_t* __restrict__ seg_begin, const int64_t* __restrict__ seg_len,
                                            const int64_t* __restrict__ tile_start, int n_pairs, int n_seg, int64_t g)
{
    int64_t p = find_le(tile_start, n_pairs, g);
    int64_t t_in = g - __ldg(tile_start + p);
    const uint64_t* base = (seg_base ? seg_base[p] : keys) + __ldg(seg_begin + p);
    SegTile s;
    s.src = base + t_in * TILE;
    s.lim = min((int64_t)TILE, __ldg(seg_len + p) - t_in * TILE);
    s.q = p / n_seg;
    return s;
}

// level-2 pass A: sizes of the sub-buckets.  Per tile a shared-memory histogram of NB2 bins, flushed with one global
// atomic per non-empty bin.
template <int TILE>
__global__ void __launch_bounds__(256)
k_sub_hist(const uint64_t* __restrict__ keys, const uint64_t* const* __restrict__ seg_base,
           const int64_t* __restrict__ seg_begin, const int64_t* __restrict__ seg_len,
           const int64_t* __restrict__ tile_start, int n_pairs, int n_seg, int pb, unsigned int NB2,
           unsigned long long* __restrict__ sub_hist)
{
    constexpr int PER = TILE / 256;
    extern __shared__ unsigned int sh_hist[];
    const int t = threadIdx.x;
    for (unsigned int i = t; i < NB2; i += 256) sh_hist[i] = 0;
    __syncthreads();
    const int64_t n_tiles = __ldg(tile_start + n_pairs);
    for (int64_t g = blockIdx.x; g < n_tiles; g += gridDim.x) {
        SegTile s = seg_tile<TILE>(keys, seg_base, seg_begin, seg_len, tile_start, n_pairs, n_seg, g);
        const uint64_t* src = s.src + t;
        const int lim = (int)s.lim - t;
#pragma unroll
        for (int j0 = 0; j0 < PER; j0 += 8) {                    // 8 loads in flight per thread, then their atomics
            uint64_t kk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) kk[j] = ((j0 + j) * 256 < lim) ? __ldcs(src + (j0 + j) * 256) : PC_KEY_SENTINEL;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (kk[j] != PC_KEY_SENTINEL) atomicAdd(&sh_hist[sub_of(hash64(kk[j]), pb, NB2)], 1u);
        }
        __syncthreads();
        for (unsigned int x = t; x < NB2; x += 256) {
            unsigned int c = sh_hist[x];
            if (c) { atomicAdd(&sub_hist[s.q * NB2 + x], (unsigned long long)c); sh_hist[x] = 0; }
        }
        __syncthreads();
    }
}

// level-2 pass B: the tile-wise counting sort of k_tile_scatter, by sub-bucket.  The source may be a peer's bucket
// buffer (NVLink): the multi-GPU exchange is fused into this read.
template <int TILE, bool PACK>
__global__ void __launch_bounds__(256, 2)
k_sub_scatter(const uint64_t* __restrict__ keys, const uint64_t* const* __restrict__ seg_base,
              const int64_t* __restrict__ seg_begin, const int64_t* __restrict__ seg_len,
              const int64_t* __restrict__ tile_start, int n_pairs, int n_seg, int pb, unsigned int NB2,
              const unsigned long long* __restrict__ sub_off, unsigned long long* __restrict__ cursor,
              uint64_t* __restrict__ out)
{
    constexpr int PER = TILE / 256;
    extern __shared__ unsigned long long sm64[];
    unsigned long long* sorted = sm64;                                   // [TILE]
    unsigned long long* gbase = sm64 + TILE;                             // [NB2]
    unsigned int* cnt = reinterpret_cast<unsigned int*>(gbase + NB2);    // [NB2]
    unsigned int* lstart = cnt + NB2;                                    // [NB2 + 1]
    unsigned short* qs = reinterpret_cast<unsigned short*>(lstart + NB2 + 1 + ((NB2 + 1) & 1));   // [TILE]
    __shared__ unsigned int wsum[8];
    __shared__ SegTile s_tile;                                           // tile descriptor (keeps it out of the registers)
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int64_t n_tiles = __ldg(tile_start + n_pairs);
    for (int64_t g = blockIdx.x; g < n_tiles; g += gridDim.x) {
        if (t == 0) s_tile = seg_tile<TILE>(keys, seg_base, seg_begin, seg_len, tile_start, n_pairs, n_seg, g);
        __syncthreads();
        uint64_t key[PER];
        {
            const uint64_t* src = s_tile.src + t;
            const int lim = (int)s_tile.lim - t;
#pragma unroll
            for (int j = 0; j < PER; ++j) key[j] = (j * 256 < lim) ? __ldcs(src + j * 256) : PC_KEY_SENTINEL;
        }
        for (unsigned int i = t; i < NB2; i += 256) cnt[i] = 0;
        __syncthreads();
        unsigned int q2[PER / 2];                                        // two 16-bit sub-bucket ids per register
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            unsigned int q = sub_of(hash64(key[j]), pb, NB2);
            if (key[j] != PC_KEY_SENTINEL) atomicAdd(&cnt[q], 1u);
            if (j & 1) q2[j >> 1] |= q << 16; else q2[j >> 1] = q;
        }
        __syncthreads();
        {
            unsigned int per = (NB2 + 255) / 256;
            unsigned int b0 = min(NB2, t * per), b1 = min(NB2, b0 + per);
            unsigned int sum = 0;
            for (unsigned int x = b0; x < b1; ++x) sum += cnt[x];
            unsigned int incl = sum;
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
            unsigned int run = before + incl - sum;
            const unsigned long long sub0 = (unsigned long long)s_tile.q * NB2;
            // the run reservations are independent returning L2 atomics: all of a thread's (up to 8) are issued back to
            // back instead of one per loop iteration (a chain of ~1 us round trips per tile once NB2 exceeds 256)
            unsigned int cc[PC_SUB_MAX / 256];
#pragma unroll
            for (int u = 0; u < PC_SUB_MAX / 256; ++u) {
                unsigned int x = b0 + u;
                cc[u] = (x < b1) ? cnt[x] : 0u;
                if (x < b1) { lstart[x] = run; cnt[x] = 0; run += cc[u]; }
            }
            unsigned long long res[PC_SUB_MAX / 256];
#pragma unroll
            for (int u = 0; u < PC_SUB_MAX / 256; ++u)
                res[u] = cc[u] ? __ldg(sub_off + sub0 + b0 + u) + atomicAdd(&cursor[sub0 + b0 + u], (unsigned long long)cc[u]) : 0ull;
#pragma unroll
            for (int u = 0; u < PC_SUB_MAX / 256; ++u)
                if (cc[u]) gbase[b0 + u] = res[u] - lstart[b0 + u];     // minus the local run start
            if (t == 255) lstart[NB2] = run;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (key[j] != PC_KEY_SENTINEL) {
                unsigned int q = (j & 1) ? (q2[j >> 1] >> 16) : (q2[j >> 1] & 0xFFFFu);
                unsigned int pos = lstart[q] + atomicAdd(&cnt[q], 1u);
                if (PACK) sorted[pos] = key[j] | ((unsigned long long)q << PC_PACK_SHIFT);
                else { sorted[pos] = key[j]; qs[pos] = (unsigned short)q; }
            }
        }
        __syncthreads();
        unsigned int nv = lstart[NB2];
        for (unsigned int i = t; i < nv; i += 256) {
            unsigned long long v = sorted[i];
            unsigned int x = PACK ? (unsigned int)(v >> PC_PACK_SHIFT) : qs[i];
            out[gbase[x] + i] = PACK ? (v & ((1ull << PC_PACK_SHIFT) - 1ull)) : v;
        }
        __syncthreads();
    }
}

// Variant of pass B with ONE shared atomic per k-mer: the returning atomicAdd of the histogram pass already is the rank
// of the k-mer inside its sub-bucket run, so the ranking pass needs no second atomic (and the counters no re-zeroing).
// (sub-bucket id, rank) are kept packed in one register per k-mer (11 + 13 bits); with 512 threads x 16 k-mers the
// register budget is the same as 256 x 32 and twice as many warps hide the latency of the returning atomics.
template <int TILE, int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
k_sub_scatter_r(const uint64_t* __restrict__ keys, const uint64_t* const* __restrict__ seg_base,
                const int64_t* __restrict__ seg_begin, const int64_t* __restrict__ seg_len,
                const int64_t* __restrict__ tile_start, int n_pairs, int n_seg, int pb, unsigned int NB2,
                const unsigned long long* __restrict__ sub_off, unsigned long long* __restrict__ cursor,
                uint64_t* __restrict__ out)
{
    constexpr int PER = TILE / THREADS;
    constexpr int WARPS = THREADS / 32;
    static_assert(TILE <= 8192 && PC_SUB_MAX <= 2048, "packing of (sub-bucket, rank) into 24 bits");
    extern __shared__ unsigned long long sm64[];
    unsigned long long* sorted = sm64;                                   // [TILE]
    unsigned long long* gbase = sm64 + TILE;                             // [NB2]
    unsigned int* cnt = reinterpret_cast<unsigned int*>(gbase + NB2);    // [NB2]
    unsigned int* lstart = cnt + NB2;                                    // [NB2 + 1]
    unsigned short* qs = reinterpret_cast<unsigned short*>(lstart + NB2 + 1 + ((NB2 + 1) & 1));   // [TILE]
    __shared__ unsigned int wsum[WARPS];
    __shared__ SegTile s_tile;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int64_t n_tiles = __ldg(tile_start + n_pairs);
    for (int64_t g = blockIdx.x; g < n_tiles; g += gridDim.x) {
        if (t == 0) s_tile = seg_tile<TILE>(keys, seg_base, seg_begin, seg_len, tile_start, n_pairs, n_seg, g);
        for (unsigned int i = t; i < NB2; i += THREADS) cnt[i] = 0;
        __syncthreads();
        uint64_t key[PER];
        {
            const uint64_t* src = s_tile.src + t;
            const int lim = (int)s_tile.lim - t;
#pragma unroll
            for (int j = 0; j < PER; ++j) key[j] = (j * THREADS < lim) ? __ldcs(src + j * THREADS) : PC_KEY_SENTINEL;
        }
        unsigned int qr[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            unsigned int q = sub_of(hash64(key[j]), pb, NB2);
            unsigned int r = (key[j] != PC_KEY_SENTINEL) ? atomicAdd(&cnt[q], 1u) : 0u;
            qr[j] = q | (r << 11);
        }
        __syncthreads();
        {
            unsigned int per = (NB2 + THREADS - 1) / THREADS;
            unsigned int b0 = min(NB2, t * per), b1 = min(NB2, b0 + per);
            unsigned int sum = 0;
            for (unsigned int x = b0; x < b1; ++x) sum += cnt[x];
            unsigned int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned int u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += u;
            }
            if (lane == 31) wsum[wid] = incl;
            __syncthreads();
            unsigned int before = 0;
#pragma unroll
            for (int x = 0; x < WARPS; ++x) if (x < wid) before += wsum[x];
            unsigned int run = before + incl - sum;
            const unsigned long long sub0 = (unsigned long long)s_tile.q * NB2;
            for (unsigned int x = b0; x < b1; ++x) {
                unsigned int c = cnt[x];
                lstart[x] = run;
                if (c) gbase[x] = sub_off[sub0 + x] + atomicAdd(&cursor[sub0 + x], (unsigned long long)c);
                run += c;
            }
            if (t == THREADS - 1) lstart[NB2] = run;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (key[j] != PC_KEY_SENTINEL) {
                unsigned int q = qr[j] & 0x7FFu;
                unsigned int pos = lstart[q] + (qr[j] >> 11);
                sorted[pos] = key[j];
                qs[pos] = (unsigned short)q;
            }
        }
        __syncthreads();
        unsigned int nv = lstart[NB2];
        for (unsigned int i = t; i < nv; i += THREADS) {
            unsigned int x = qs[i];
            out[gbase[x] + (i - lstart[x])] = sorted[i];
        }
        __syncthreads();
    }
}

// level-2 pass C: one CTA per sub-bucket at a time (persistent, strided: neighbouring CTAs stream neighbouring
// sub-buckets).  Shared-memory table: keys[SLOTS] (stored complemented, 0 = empty) + counts[SLOTS] (occurrences - 1: the
// CAS that claims a slot records the first occurrence, so a new k-mer costs one shared atomic, a repeat two).
// The k-mers of the next batch / next sub-bucket are requested from HBM before the current batch is inserted.
// Output: (k-mer, count) of every k-mer with count >= keep_min, appended to a list with one global atomic per
// sub-bucket; ids of sub-buckets that did not fit go to ovf_ids.
// The kernel is instruction-issue bound (ncu: 63 % issue active, DRAM 11 %), so the per-sub-bucket epilogue is written
// for few instructions: every lane owns 4 consecutive slots per step (one 128-bit load of the counts, three 128-bit
// stores to re-zero them), and the slot hash is one 32-bit multiply -- the sub-bucket was chosen by the strong 64-bit
// mix, so inside it any cheap independent hash spreads the k-mers.
#define PC_SMEM_PROBE_LIMIT 1024u
template <int SLOTS>
__device__ __forceinline__ unsigned int smem_slot(uint64_t key) {
    constexpr int LOG = SLOTS == 4096 ? 12 : SLOTS == 8192 ? 13 : SLOTS == 16384 ? 14 : -1;
    static_assert(LOG > 0, "table size");
    return (((unsigned int)key ^ (unsigned int)(key >> 32)) * 0x9E3779B1u) >> (32 - LOG);
}

template <int SLOTS, int THREADS, int PER, bool PAIR>
__global__ void __launch_bounds__(THREADS, (2048 / THREADS) < (int)((220u << 10) / (SLOTS * 12u + 1024u)) ? (2048 / THREADS) : (int)((220u << 10) / (SLOTS * 12u + 1024u)))
k_count_smem(const uint64_t* __restrict__ kbuf2, const unsigned long long* __restrict__ sub_off, unsigned int n_sub,
             unsigned int keep_min, uint64_t* __restrict__ list_keys, uint32_t* __restrict__ list_counts,
             unsigned long long list_cap, unsigned long long* __restrict__ n_list,
             unsigned int* __restrict__ ovf_ids, unsigned int* __restrict__ n_ovf, CountStats* __restrict__ stats)
{
    constexpr int WARPS = THREADS / 32;
    constexpr int SPW = SLOTS / WARPS;                           // slots scanned by one warp
    constexpr int IT4 = SPW / 128;                               // steps of 32 lanes x 4 slots
    constexpr unsigned int BATCH = THREADS * PER;
    static_assert(SLOTS % (WARPS * 128) == 0 && IT4 * 4 <= 32, "slot layout");
    extern __shared__ __align__(16) unsigned long long skeys[];  // [SLOTS]
    unsigned int* scnt = reinterpret_cast<unsigned int*>(skeys + SLOTS);   // [SLOTS]
    __shared__ unsigned int s_total, s_ovf;
    __shared__ unsigned long long s_gbase;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const unsigned int G = gridDim.x;

    for (int i = t; i < SLOTS; i += THREADS) { skeys[i] = 0ull; scnt[i] = 0u; }
    if (t == 0) { s_total = 0; s_ovf = 0; }
    __syncthreads();

    unsigned long long distinct = 0;
    unsigned int max_probe = 0;
    // pipeline state: (cur_sb, cur_i0, cur_end) describe the batch whose k-mers are in nkey[]
    unsigned int cur_sb = blockIdx.x;
    if (cur_sb >= n_sub) return;
    unsigned long long cur_i0 = __ldg(sub_off + cur_sb), cur_end = __ldg(sub_off + cur_sb + 1);
    unsigned long long nbeg = 0, nend = 0;                       // bounds of sub-bucket cur_sb + G
    if (cur_sb + G < n_sub) { nbeg = __ldg(sub_off + cur_sb + G); nend = __ldg(sub_off + cur_sb + G + 1); }
    uint64_t nkey[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        unsigned long long i = cur_i0 + (unsigned long long)j * THREADS + t;
        nkey[j] = (i < cur_end) ? __ldcs(kbuf2 + i) : PC_KEY_SENTINEL;
    }
    unsigned int nk = 0;                                         // slots this thread claimed in the current sub-bucket
    for (;;) {
        uint64_t key[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) key[j] = nkey[j];
        const unsigned int this_sb = cur_sb;
        const bool last_batch = cur_i0 + BATCH >= cur_end;
        if (!last_batch) cur_i0 += BATCH;
        else {
            cur_sb += G; cur_i0 = nbeg; cur_end = nend;
            if (cur_sb + G < n_sub) { nbeg = __ldg(sub_off + cur_sb + G); nend = __ldg(sub_off + cur_sb + G + 1); }
        }
        const bool more = cur_sb < n_sub;
        if (more) {
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                unsigned long long i = cur_i0 + (unsigned long long)j * THREADS + t;
                nkey[j] = (i < cur_end) ? __ldcs(kbuf2 + i) : PC_KEY_SENTINEL;
            }
        }
        // ---- insert this batch
        if (PAIR && *(volatile unsigned int*)&s_ovf == 0u) {
            // Experiment kept as smem_variant 4: read-first insertion, two slots per probe step (one 128-bit shared load
            // shows an aligned slot pair; a repeat costs that load + one atomicAdd, a new k-mer the load + one CAS, only
            // a pair taken by two OTHER k-mers sends the lane round the loop).  Meant to cut the divergent probe loops
            // of CAS-first linear probing (68 % of this kernel's instructions); MEASURED SLOWER, 6.1 vs 4.0 ms at
            // 4.96e8 k-mers: the random 16-byte shared loads and the longer decision chain cost more than the loops.
            // Slots only ever change from empty to a k-mer, so a non-empty read is final and a stale empty read is
            // settled by the CAS.
            ulonglong2 kk[PER]; unsigned int slot[PER];
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                slot[j] = smem_slot<SLOTS>(key[j]) & ~1u;
                kk[j] = *reinterpret_cast<const ulonglong2*>(skeys + slot[j]);
            }
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                if (key[j] == PC_KEY_SENTINEL) continue;
                const unsigned long long skey = ~key[j];
                ulonglong2 cur = kk[j];
                unsigned int b = slot[j], probes = 0;
                for (;;) {
                    int target = -1; bool fresh = false;
                    if (cur.x == skey) target = (int)b;
                    else if (cur.x == 0ull) { target = (int)b; fresh = true; }
                    else if (cur.y == skey) target = (int)b + 1;
                    else if (cur.y == 0ull) { target = (int)b + 1; fresh = true; }
                    if (target >= 0) {
                        if (!fresh) { atomicAdd(&scnt[target], 1u); break; }
                        unsigned long long o = atomicCAS(&skeys[target], 0ull, skey);
                        if (o == 0ull) { ++nk; break; }
                        if (o == skey) { atomicAdd(&scnt[target], 1u); break; }
                        if (target == (int)b) cur.x = o; else cur.y = o;   // lost the slot to another k-mer: look again
                        continue;
                    }
                    if (++probes > PC_SMEM_PROBE_LIMIT) { s_ovf = 1u; break; }
                    b = (b + 2) & (SLOTS - 1);
                    cur = *reinterpret_cast<const ulonglong2*>(skeys + b);
                }
                max_probe = max(max_probe, probes);
            }
        }
        if (!PAIR && *(volatile unsigned int*)&s_ovf == 0u) {
            unsigned long long old[PER]; unsigned int slot[PER];
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                slot[j] = smem_slot<SLOTS>(key[j]);
                old[j] = (key[j] != PC_KEY_SENTINEL) ? atomicCAS(&skeys[slot[j]], 0ull, ~key[j]) : 0ull;
            }
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                if (key[j] == PC_KEY_SENTINEL) continue;
                const unsigned long long skey = ~key[j];
                unsigned long long o = old[j];
                unsigned int s = slot[j], probes = 0;
                for (;;) {
                    if (o == 0ull) { ++nk; break; }
                    if (o == skey) { atomicAdd(&scnt[s], 1u); break; }
                    if (++probes > PC_SMEM_PROBE_LIMIT) { s_ovf = 1u; break; }
                    s = (s + 1) & (SLOTS - 1);
                    o = atomicCAS(&skeys[s], 0ull, skey);
                }
                max_probe = max(max_probe, probes);
            }
        }
        if (!last_batch) continue;

        // ---- sub-bucket complete: emit the k-mers that reach keep_min, re-zero the table
        __syncthreads();                                         // A: all inserts of this_sb done
        const bool ovf = s_ovf != 0u;
        if (!ovf) distinct += nk;
        nk = 0;
        uint4 c4[IT4];
        unsigned int emask = 0, n_emit = 0;                      // bit (4 i + j): slot j of step i is emitted
        const int s0 = wid * SPW + lane * 4;
#pragma unroll
        for (int i = 0; i < IT4; ++i) {
            const int s = s0 + i * 128;
            c4[i] = *reinterpret_cast<const uint4*>(scnt + s);
            const unsigned int cc[4] = {c4[i].x, c4[i].y, c4[i].z, c4[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                bool e = !ovf && cc[j] + 1u >= keep_min;
                if (e && keep_min <= 1u) e = skeys[s + j] != 0ull;   // count field 0 is also what an empty slot holds
                emask |= (unsigned int)e << (4 * i + j);
                n_emit += e;
            }
        }
        unsigned int incl = n_emit;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned int u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += u;
        }
        const unsigned int wtot = __shfl_sync(0xFFFFFFFFu, incl, 31);
        unsigned int wbase = 0;
        if (lane == 0 && wtot) wbase = atomicAdd(&s_total, wtot);
        wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
        __syncthreads();                                         // B: s_total complete, everybody has read s_ovf
        if (t == 0) {
            unsigned int tot = s_total;
            s_gbase = tot ? atomicAdd(n_list, (unsigned long long)tot) : 0ull;
            s_total = 0; s_ovf = 0;
            if (ovf) ovf_ids[atomicAdd(n_ovf, 1u)] = this_sb;
        }
        __syncthreads();                                         // C: s_gbase valid
        unsigned long long pos = s_gbase + wbase + (incl - n_emit);
#pragma unroll
        for (int i = 0; i < IT4; ++i) {
            const int s = s0 + i * 128;
            if ((emask >> (4 * i)) & 0xFu) {
                const unsigned int cc[4] = {c4[i].x, c4[i].y, c4[i].z, c4[i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if ((emask >> (4 * i + j)) & 1u) {
                        if (pos < list_cap) { list_keys[pos] = ~skeys[s + j]; list_counts[pos] = cc[j] + 1u; }
                        ++pos;
                    }
            }
            __syncwarp();                                        // the key stores below cover other lanes' slots
            *reinterpret_cast<uint4*>(scnt + s) = make_uint4(0u, 0u, 0u, 0u);
            uint4* kz = reinterpret_cast<uint4*>(skeys + wid * SPW + i * 128);   // 128 slots = 64 x 16 bytes
            kz[lane] = make_uint4(0u, 0u, 0u, 0u);
            kz[lane + 32] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();                                         // D: table is empty again
        if (!more) break;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        distinct += __shfl_xor_sync(0xFFFFFFFFu, distinct, o);
        max_probe = max(max_probe, __shfl_xor_sync(0xFFFFFFFFu, max_probe, o));
    }
    if (lane == 0) {
        if (distinct) atomicAdd(&stats->distinct, distinct);
        if (max_probe > 8) atomicMax(&stats->max_probe, (unsigned long long)max_probe);
    }
}

// largest counting unit: max over i of off[i + 1] - off[i] (the uint32 counters of a unit cannot wrap while the unit holds
// fewer than 2^32 k-mer instances; the host refuses to return counts otherwise)
__global__ void __launch_bounds__(256)
k_max_diff(const unsigned long long* __restrict__ off, int64_t n, unsigned long long* __restrict__ out)
{
    unsigned long long m = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = max(m, off[i + 1] - off[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// K4 over the compact list: survivors appended with one atomic per warp (same contract as k_select)
__global__ void __launch_bounds__(256)
k_select_list(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts, int64_t n, unsigned int low,
              unsigned int high, uint64_t* __restrict__ out_keys, uint32_t* __restrict__ out_counts,
              unsigned long long out_cap, unsigned long long* __restrict__ n_out)
{
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t n_round = (n + 31) & ~(int64_t)31;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        unsigned int c = (i < n) ? __ldcs(counts + i) : 0u;
        bool keep = i < n && c >= low && c <= high;
        unsigned int bal = __ballot_sync(0xFFFFFFFFu, keep);
        if (bal) {
            int lane = threadIdx.x & 31;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(n_out, (unsigned long long)__popc(bal));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (keep && out_keys) {
                unsigned long long pos = base + __popc(bal & ((1u << lane) - 1u));
                if (pos < out_cap) {
                    out_keys[pos] = __ldcs(keys + i);
                    if (out_counts) out_counts[pos] = c;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ K5, filtered probe
// The stream probe is bound by random L2 accesses to the repeat table (ncu: DRAM 10 %, every warp on long-scoreboard;
// half of the accesses cross the die-to-die fabric).  Most windows are NOT repeat k-mers, so a 1-bit-per-k-mer filter
// held in SHARED memory (up to ~1.7 Mbit: one hash, load ~0.4) answers "certainly not a repeat" for about two thirds of
// them without leaving the SM; only the rest goes to L2.  Persistent CTAs (one per SM, the filter fills its shared
// memory) load the bitmap once and then walk word tiles.  Used when the repeat set is small enough for the filter to
// discriminate (host side: bits >= 2 x n_selected).
__device__ __forceinline__ unsigned int filt_bit(uint64_t key, unsigned int nbits) { return gbit_of(key, nbits); }

__global__ void __launch_bounds__(256)
k_filter_build(const uint64_t* __restrict__ sorted_keys, int64_t n, unsigned int nbits, unsigned int* __restrict__ bitmap)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned int b = filt_bit(sorted_keys[i], nbits);
    atomicOr(&bitmap[b >> 5], 1u << (b & 31));
}

template <int BATCH, int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
k_probe_stream_filt(const uint64_t* __restrict__ stream, int64_t n_words,
                    const int64_t* __restrict__ chunk_word0, int64_t n_chunks, const int32_t* __restrict__ chunk_row,
                    const int64_t* __restrict__ row_hit_base, const RepSlot* __restrict__ rtable, uint64_t rcap,
                    const unsigned int* __restrict__ bitmap, unsigned int nbits,
                    int32_t* __restrict__ hits, unsigned int* __restrict__ row_cursor)
{
    extern __shared__ __align__(16) unsigned int sbits[];        // [(nbits + 31) / 32 rounded up to 4]
    {
        const unsigned int n4 = ((nbits + 31) / 32 + 3) / 4;
        const uint4* src = reinterpret_cast<const uint4*>(bitmap);
        uint4* dst = reinterpret_cast<uint4*>(sbits);
        for (unsigned int i = threadIdx.x; i < n4; i += THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t n_tiles = (n_words + THREADS - 1) / THREADS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int64_t w = tile * THREADS + threadIdx.x;
        int64_t w_lane0 = w - lane;
        if (w_lane0 >= n_words) continue;                        // whole warp past the end
        int row = -1;
        if (w < n_words) {
            int64_t c = chunk_of_word(chunk_word0, n_chunks, w, w_lane0);
            row = __ldg(chunk_row + c);
        }
        bool lane_on = row >= 0;
        unsigned int on_mask = __ballot_sync(0xFFFFFFFFu, lane_on);
        if (on_mask == 0u) continue;
        int row0 = __shfl_sync(0xFFFFFFFFu, row, __ffs(on_mask) - 1);
        bool uniform = __all_sync(0xFFFFFFFFu, !lane_on || row == row0);
        const uint64_t* src = stream + w_lane0 * 32 + lane;
        // software pipeline: the stream k-mers of the next batch are requested (DRAM) before this batch's table probes
        // (L2) are issued -- ncu showed a third of the stall samples on the first use of the stream loads
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
                    unsigned int fb = filt_bit(key, nbits);
                    bool maybe = key != PC_KEY_SENTINEL && ((sbits[fb >> 5] >> (fb & 31)) & 1u);
                    vmask |= (unsigned int)maybe << j;
                    skey[j] = ~key;
                    slot[j] = rep_slot(key, rcap);
                }
#pragma unroll
                for (int j = 0; j < BATCH; ++j)
                    cur[j] = ((vmask >> j) & 1u) ? __ldg(reinterpret_cast<const uint4*>(rtable + slot[j])) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    cols[j] = 0;
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
}

// ------------------------------------------------------------------------------------------------ SM-driven copies
// Host <-> device copies issued by a few CTAs instead of a copy engine (the host side is page-locked memory, which
// under unified addressing is mapped into the device's address space).  Every stage of the hot path reads a few
// counters back and uploads small tables with tiny copies on the copy engines; queued behind a 400 MB transfer of a
// neighbouring pipeline step on the same engine each of them waits milliseconds (measured: the overlapped end-to-end
// loop ran at 40 ms per step instead of ~21 ms, the engines serve their queue in order).  Copies made by SMs leave the
// engines to the small ones.  16 bytes per thread, 4 independent transfers in flight per thread.
__global__ void __launch_bounds__(256)
k_copy16(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t n16)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < n16; i += stride) dst[i] = src[i];
}
__global__ void __launch_bounds__(32)
k_copy_tail(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n)
{
    if ((int)threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
}

// ------------------------------------------------------------------------------------------------ N2 diagnostics
// kcount_hist (polycracker.py:1419-1496): Counter over the counts of the dumped k-mers -> how many distinct k-mers occur
// exactly c times.  hist[c - lo] for lo <= c <= hi, hist[hi - lo + 1] collects everything above hi.
__global__ void __launch_bounds__(256)
k_count_hist_list(const uint32_t* __restrict__ counts, int64_t n, unsigned int lo, unsigned int hi,
                  unsigned long long* __restrict__ hist)
{
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        unsigned int c = __ldcs(counts + i);
        if (c >= lo) atomicAdd(&hist[min(c, hi + 1u) - lo], 1ull);
    }
}
__global__ void __launch_bounds__(256)
k_count_hist_table(CountTable table, uint64_t capacity, unsigned int lo, unsigned int hi, unsigned long long* __restrict__ hist)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < capacity; i += stride) {
        if (__ldcs(table.keys + i) == 0ull) continue;
        unsigned int c = __ldcs(table.counts + i) + 1u;           // stored = occurrences - 1
        if (c >= lo) atomicAdd(&hist[min(c, hi + 1u) - lo], 1ull);
    }
}
// number_repeatmers_per_subsequence (polycracker.py:1366-1382): k-mer names on a chunk's blasted_merged.bed line = exact
// occurrences of repeat k-mers in the chunk = what the probe appended to the row
__global__ void __launch_bounds__(256)
k_row_hits(const unsigned int* __restrict__ row_cursor, int64_t n_rows, int64_t* __restrict__ out)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_rows) out[r] = (int64_t)row_cursor[r];
}

// ------------------------------------------------------------------------------------------------ N3 standard scaling
// The other normalisation at the boundary (transform_plot with tfidf = 0): generate_Kmer_Matrix's length normalisation
// (polycracker.py:357-359: row i times 1 / (len_i / 5000.), stored as float32(float64(v) * (1/x))) followed by sklearn's
// StandardScaler(with_mean=False) (polycracker.py:392-393): column j divided by its standard deviation over ALL rows
// (zeros included).  sklearn (1.9, sparsefuncs_fast.pyx, float32 input): means and variances accumulated in double,
// variance = (sum over stored entries of (x - mean)^2 + (n - nnz_j) mean^2) / n, both rounded to float32 and widened
// again, near-constant columns (var <= n eps var + (n mean eps)^2) get scale 1, values multiplied by 1/sqrt(var) in
// double and rounded to float32 once.
__device__ __forceinline__ float scaled_value(float v, const double* __restrict__ row_scale, int64_t row) {
    return row_scale ? (float)((double)v * row_scale[row]) : v;
}
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_col_sum(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ data,
          const double* __restrict__ row_scale, double* __restrict__ col_sum)
{
    int64_t row = blockIdx.x;
    for (int64_t i = indptr[row] + threadIdx.x; i < indptr[row + 1]; i += THREADS)
        atomicAdd(&col_sum[indices[i]], (double)scaled_value(data[i], row_scale, row));
}
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_col_sqdev(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ data,
            const double* __restrict__ row_scale, const double* __restrict__ col_sum, double n_rows,
            double* __restrict__ col_ss)
{
    int64_t row = blockIdx.x;
    for (int64_t i = indptr[row] + threadIdx.x; i < indptr[row + 1]; i += THREADS) {
        int c = indices[i];
        double d = (double)scaled_value(data[i], row_scale, row) - col_sum[c] / n_rows;
        atomicAdd(&col_ss[c], d * d);
    }
}
__global__ void __launch_bounds__(256)
k_col_factor(const double* __restrict__ col_sum, const double* __restrict__ col_ss, const int* __restrict__ df,
             int64_t n_cols, double n_rows, double* __restrict__ factor)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_cols) return;
    double mean = col_sum[j] / n_rows;
    double var = (col_ss[j] + (n_rows - (double)df[j]) * mean * mean) / n_rows;
    double mean32 = (double)(float)mean, var32 = (double)(float)var;       // mean_variance_axis returns float32 for float32 input
    const double eps = 2.220446049250313e-16;
    double bound = n_rows * eps * var32 + (n_rows * mean32 * eps) * (n_rows * mean32 * eps);
    double scale = (var32 <= bound) ? 1.0 : sqrt(var32);
    factor[j] = 1.0 / scale;
}
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_col_apply(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ data,
            const double* __restrict__ row_scale, const double* __restrict__ factor, float* __restrict__ out)
{
    int64_t row = blockIdx.x;
    for (int64_t i = indptr[row] + threadIdx.x; i < indptr[row + 1]; i += THREADS)
        out[i] = (float)((double)scaled_value(data[i], row_scale, row) * factor[indices[i]]);
}

// ------------------------------------------------------------------------------------------------ N3 Gram matrix, panels
// X X^T of the (TF-IDF / scaled / count) matrix for the linear and cosine kernels of KernelPCA (polycracker.py:387, 398-399)
// is the first dense contraction behind the hot path.  X is R x D with ~2 % of its entries stored; the contraction runs
// over column panels that are densified on the fly: row r of the panel = the entries of CSR row r with col0 <= column <
// col0 + ncols (a binary search finds the first), everything else 0.  One CTA per row.
template <typename T>
__device__ __forceinline__ T gram_cast(float v);
template <> __device__ __forceinline__ float gram_cast<float>(float v) { return v; }

template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_densify_panel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
                int64_t col0, int ncols, T* __restrict__ out, int64_t ld, int round_tf32 /* float panels for the tensor-core Gram kernel */)
{
    const int64_t row = blockIdx.x;
    T* o = out + row * ld;
    for (int i = threadIdx.x; i < ncols; i += THREADS) o[i] = gram_cast<T>(0.0f);
    __syncthreads();
    const int64_t beg = indptr[row], end = indptr[row + 1];
    int64_t lo = beg, hi = end;                               // first entry with column >= col0
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if ((int64_t)indices[mid] < col0) lo = mid + 1; else hi = mid; }
    for (int64_t i = lo + threadIdx.x; i < end; i += THREADS) {
        const int64_t c = (int64_t)indices[i] - col0;
        if (c >= ncols) break;                                // columns ascend inside a row
        float v = values[i];
        if (round_tf32) {                                       // round to nearest TF32 here instead of in a pass of its own
            uint32_t r;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
            v = __uint_as_float(r);
        }
        o[c] = gram_cast<T>(v);
    }
}

}  // namespace pc
