// K3, super-k-mer path (part_mode 3): packed genome -> super-k-mer records -> two-level radix partition of the RECORDS ->
// shared-memory count that expands the k-mers of a record in registers.  See kmer_skm.cuh for the record format and
// why the partition stays exact.  Compared with the k-mer path (kernels_smem.cuh) the partition passes move ~2 bytes
// per window instead of 8 and pay their shared-memory atomics per record (6-7 windows) instead of per window.
#pragma once
#include "kernels_smem.cuh"
#include "kmer_skm.cuh"

namespace pc {

typedef ulonglong2 SkmRec;                                     // .x = hi (first 32 bases), .y = lo (bases, dest, n - 1)

// ------------------------------------------------------------------------------------------------ pass A: extract
// One thread = one packed word = 32 window starts (the k - 1 bases of overlap come from the next lane by shuffle).
// Phase 1 (fully unrolled, registers): window validity, rolling canonical m-mer hashes, sliding minimum, record
// boundaries.  Phase 2 (a short loop over the ~5 records of the thread): destination from the minimizer, record
// assembly with two funnel shifts, one global RED for the sub-bucket histogram, one shared atomic for a slot in the
// CTA's staging area.  The CTA then reserves its output range with ONE global atomic and copies the staged records
// out fully coalesced.  Records beyond the staging capacity (very short runs) go out one by one.
#define PC_SKM_THREADS 256
template <int W>
__global__ void __launch_bounds__(PC_SKM_THREADS)
k_skm_extract(const uint64_t* __restrict__ words, const uint32_t* __restrict__ nmask, int64_t n_words, int k, int m, int nmax,
              int pb, unsigned int nb2, SkmRec* __restrict__ out, unsigned long long out_cap,
              unsigned long long* __restrict__ n_out /* [0] records, [1] set when out_cap was exceeded */,
              unsigned long long* __restrict__ sub_hist, unsigned int stage_cap, CountStats* __restrict__ stats)
{
    extern __shared__ __align__(16) unsigned char skm_smem[];
    uint32_t* sh_h = reinterpret_cast<uint32_t*>(skm_smem);                                  // [32][THREADS]
    SkmRec* stage = reinterpret_cast<SkmRec*>(skm_smem + 32 * PC_SKM_THREADS * sizeof(uint32_t));   // [stage_cap]
    __shared__ unsigned int s_n;
    __shared__ unsigned long long s_base;
    const int t = threadIdx.x, lane = t & 31;
    if (t == 0) s_n = 0;
    __syncthreads();
    const int64_t w = (int64_t)blockIdx.x * PC_SKM_THREADS + t;
    WordPair p = load_word_pair(words, nmask, w, n_words);
    uint32_t vm = 0, st = 0;
    if (p.m0 != 0xFFFFFFFFu) {
        SkmThread<W> th;
        th.run(p.w0, p.w1, p.m0, p.m1, k, m, nmax);
        vm = th.vmask; st = th.starts;
#pragma unroll
        for (int b = 0; b < 32; ++b) sh_h[b * PC_SKM_THREADS + t] = th.hmin[b];
    }
    uint32_t rem = st;
    while (rem) {
        const int b = __ffs(rem) - 1;
        rem &= rem - 1;
        const int n = skm_run_length(vm, st, b);
        const uint32_t dest = skm_dest(sh_h[b * PC_SKM_THREADS + t], pb, nb2);
        uint64_t hi, lo;
        skm_record(p.w0, p.w1, b, n, k, dest, hi, lo);
        // one 64-bit RED carries both counts of the sub-bucket: k-mer instances in the high half, records in the low half
        atomicAdd(&sub_hist[(uint64_t)(dest >> PC_SKM_SUB_BITS) * nb2 + (dest & ((1u << PC_SKM_SUB_BITS) - 1u))],
                  ((unsigned long long)n << 32) | 1ull);
        const unsigned int pos = atomicAdd(&s_n, 1u);
        if (pos < stage_cap) stage[pos] = make_ulonglong2(hi, lo);
        else {
            const unsigned long long g = atomicAdd(n_out, 1ull);
            if (g < out_cap) out[g] = make_ulonglong2(hi, lo); else n_out[1] = 1ull;
        }
    }
    __syncthreads();
    const unsigned int tot = min(s_n, stage_cap);
    if (t == 0) s_base = tot ? atomicAdd(n_out, (unsigned long long)tot) : 0ull;
    __syncthreads();
    const unsigned long long base = s_base;
    for (unsigned int i = t; i < tot; i += PC_SKM_THREADS) {
        const unsigned long long g = base + i;
        if (g < out_cap) out[g] = stage[i]; else n_out[1] = 1ull;
    }
    unsigned int nvalid = __popc(vm);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(0xFFFFFFFFu, nvalid, o);
    if (lane == 0 && nvalid) atomicAdd(&stats->valid_windows, (unsigned long long)nvalid);
}

// ------------------------------------------------------------------------------------------------ pass B: record scatter
// The tile-wise counting sort of k_tile_scatter / k_sub_scatter on 16-byte records; the key is read out of the record.
// LEVEL 1: the flat record stream (n_flat records, a device counter) -> level-1 buckets; key = dest >> 11, the run of
//          bucket q starts at off[q * nb2] (the exclusive scan of the sub-bucket histogram gives both levels' offsets).
// LEVEL 2: tiles of ONE level-1 bucket (a tile never spans two buckets or two sender segments; the source of a tile may
//          be a peer's record buffer over NVLink) -> sub-buckets; key = dest & 2047, run start off[q * nb2 + key].
template <int TILE, int LEVEL>
__global__ void __launch_bounds__(256, TILE <= 2048 ? 3 : 2)
k_rec_scatter(const SkmRec* __restrict__ recs, const SkmRec* const* __restrict__ seg_base,
              const int64_t* __restrict__ seg_begin, const int64_t* __restrict__ seg_len,
              const int64_t* __restrict__ tile_start, int n_pairs, int n_seg,
              const unsigned long long* __restrict__ n_flat, unsigned long long flat_cap, unsigned int F /* fan-out */, unsigned int nb2,
              const unsigned long long* __restrict__ off, unsigned long long* __restrict__ cursor,
              SkmRec* __restrict__ out)
{
    constexpr int PER = TILE / 256;
    extern __shared__ __align__(16) unsigned char rs_raw[];
    SkmRec* sorted = reinterpret_cast<SkmRec*>(rs_raw);                                      // [TILE]
    unsigned long long* gbase = reinterpret_cast<unsigned long long*>(sorted + TILE);        // [F]
    unsigned int* cnt = reinterpret_cast<unsigned int*>(gbase + F);                          // [F]
    unsigned int* lstart = cnt + F;                                                          // [F + 1]
    unsigned short* qs = reinterpret_cast<unsigned short*>(lstart + F + 1 + ((F + 1) & 1));  // [TILE]
    __shared__ unsigned int wsum[8];
    __shared__ const SkmRec* s_src;
    __shared__ long long s_lim, s_q;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    int64_t n_tiles;
    const unsigned long long flat_n = LEVEL == 1 ? min(*n_flat, flat_cap) : 0ull;   // (an overflowed stream is reported by the host)
    if (LEVEL == 1) n_tiles = (int64_t)((flat_n + TILE - 1) / TILE);
    else n_tiles = __ldg(tile_start + n_pairs);
    for (int64_t g = blockIdx.x; g < n_tiles; g += gridDim.x) {
        if (t == 0) {
            if (LEVEL == 1) {
                s_src = recs + g * TILE;
                s_lim = (long long)min((unsigned long long)TILE, flat_n - (unsigned long long)(g * TILE));
                s_q = 0;
            } else {
                int64_t pr = find_le(tile_start, n_pairs, g);
                int64_t t_in = g - __ldg(tile_start + pr);
                const SkmRec* base = (seg_base ? seg_base[pr] : recs) + __ldg(seg_begin + pr);
                s_src = base + t_in * TILE;
                s_lim = min((int64_t)TILE, __ldg(seg_len + pr) - t_in * TILE);
                s_q = pr / n_seg;
            }
        }
        for (unsigned int i = t; i < F; i += 256) cnt[i] = 0;
        __syncthreads();
        SkmRec rec[PER];
        unsigned short key[PER];
        {
            const SkmRec* src = s_src + t;
            const int lim = (int)s_lim - t;
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                if (j * 256 < lim) {
                    const uint4 v = ldnc_u4(src + j * 256);
                    rec[j].x = ((unsigned long long)v.y << 32) | v.x;
                    rec[j].y = ((unsigned long long)v.w << 32) | v.z;
                } else rec[j] = make_ulonglong2(0ull, 0ull);
            }
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                const uint32_t dest = skm_rec_dest(rec[j].y);
                key[j] = (unsigned short)(LEVEL == 1 ? (dest >> PC_SKM_SUB_BITS) : (dest & ((1u << PC_SKM_SUB_BITS) - 1u)));
                if (j * 256 < lim) atomicAdd(&cnt[key[j]], 1u);
                else key[j] = 0xFFFFu;
            }
        }
        __syncthreads();
        {
            const unsigned int per = (F + 255) / 256;
            const unsigned int b0 = min(F, t * per), b1 = min(F, b0 + per);
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
            const unsigned long long q = (unsigned long long)s_q;
            for (unsigned int x = b0; x < b1; ++x) {
                const unsigned int c = cnt[x];
                lstart[x] = run;
                if (c) {
                    const unsigned long long idx = LEVEL == 1 ? (unsigned long long)x * nb2 : q * nb2 + x;
                    gbase[x] = __ldg(off + idx) + atomicAdd(&cursor[LEVEL == 1 ? (unsigned long long)x : idx], (unsigned long long)c) - run;
                }
                cnt[x] = 0;
                run += c;
            }
            if (t == 255) lstart[F] = run;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (key[j] != 0xFFFFu) {
                const unsigned int pos = lstart[key[j]] + atomicAdd(&cnt[key[j]], 1u);
                sorted[pos] = rec[j];
                qs[pos] = key[j];
            }
        }
        __syncthreads();
        const unsigned int nv = lstart[F];
        for (unsigned int i = t; i < nv; i += 256) out[gbase[qs[i]] + i] = sorted[i];
        __syncthreads();
    }
}

// per level-1 bucket: number of tiles and the (begin, length) of its single local segment, from the scanned histogram
// (single GPU: no host round trip between the level-1 scatter and the level-2 scatter)
__global__ void __launch_bounds__(1024)
k_skm_tile_plan(const unsigned long long* __restrict__ sub_off, unsigned int P, unsigned int nb2, int tile,
                int64_t* __restrict__ seg_begin, int64_t* __restrict__ seg_len, int64_t* __restrict__ tile_start)
{
    __shared__ long long part[1024];
    const int t = threadIdx.x;
    const unsigned int per = (P + 1023) / 1024;
    const unsigned int beg = min(P, t * per), end = min(P, beg + per);
    long long s = 0;
    for (unsigned int q = beg; q < end; ++q) {
        const long long b = (long long)sub_off[(unsigned long long)q * nb2], e = (long long)sub_off[(unsigned long long)(q + 1) * nb2];
        seg_begin[q] = b; seg_len[q] = e - b;
        s += (e - b + tile - 1) / tile;
    }
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        long long v = (t >= o) ? part[t - o] : 0ll;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    long long run = part[t] - s;
    for (unsigned int q = beg; q < end; ++q) { tile_start[q] = run; run += (seg_len[q] + tile - 1) / tile; }
    if (t == 1023) tile_start[P] = part[1023];
}

// ------------------------------------------------------------------------------------------------ pass C: count
// k_count_smem with the k-mer load replaced by a record load + expansion.  One thread loads one record (16 bytes,
// coalesced; the record of the next batch / next sub-bucket is requested from HBM before the current one is expanded),
// but a record holds 1..nmax k-mers, and a lane that rolled through ITS record would leave most of the warp idle behind
// the longest one (measured: 9.7 ms instead of 4 for 4.96e8 k-mers, lane utilisation ~35 %).  So the k-mers of a warp's 32
// records are spread evenly over the lanes: a warp scan of the record lengths numbers the k-mers, a 512-byte map per
// warp (shared memory) says which lane holds the record of k-mer g, the record comes over by shuffle and the k-mer is
// cut out of it directly (two funnel shifts, BREV-based reverse complement).  GRP k-mers per lane are in flight: their
// CAS are issued back to back before any result is looked at.
// Insertion and epilogue as in k_count_smem: keys complemented (0 = empty), counts = occurrences - 1, CAS first, then
// linear probing; k-mers with count >= keep_min are appended to the (k-mer, count) list with one global atomic per
// sub-bucket, the table is re-zeroed; a sub-bucket that does not fit is listed in ovf_ids and re-counted through a
// global table by k_count_rec_table.
#define PC_SKM_COUNT_NMAX 16                                   // records of at most 16 k-mers: 32 x 16 map entries per warp
// a probe sequence this long means the table is > 90 % full: give the sub-bucket to the global-table fall-back now instead of
// crawling through the last slots (minimizer buckets overflow for real: heavy minimizers of repeat families)
#define PC_SKM_PROBE_LIMIT 128u
#define PC_SKM_NCAND 2048u                                     // candidate slots per sub-bucket (k-mers that reached keep_min)
template <int SLOTS, int THREADS, int GRP>
__global__ void __launch_bounds__(THREADS, (2048 / THREADS) < (int)((220u << 10) / (SLOTS * 12u + THREADS * 16u + PC_SKM_NCAND * 2u + 1024u)) ? (2048 / THREADS) : (int)((220u << 10) / (SLOTS * 12u + THREADS * 16u + PC_SKM_NCAND * 2u + 1024u)))
k_count_skm(const SkmRec* __restrict__ recs, const unsigned long long* __restrict__ sub_off, unsigned int n_sub, int k,
            unsigned int keep_min, uint64_t* __restrict__ list_keys, uint32_t* __restrict__ list_counts,
            unsigned long long list_cap, unsigned long long* __restrict__ n_list,
            unsigned int* __restrict__ ovf_ids, unsigned int* __restrict__ n_ovf, CountStats* __restrict__ stats,
            unsigned int* __restrict__ queue /* zeroed: next sub-bucket to hand out */)
{
    constexpr int WARPS = THREADS / 32;
    constexpr int SPW = SLOTS / WARPS;                           // slots owned (scanned, emitted, re-zeroed) by one warp
    constexpr int IT4 = SPW / 128;
    static_assert(SLOTS % (WARPS * 128) == 0 && IT4 * 4 <= 32 && SLOTS <= 65536, "slot layout");
    extern __shared__ __align__(16) unsigned long long skeys[];  // [SLOTS]
    unsigned int* scnt = reinterpret_cast<unsigned int*>(skeys + SLOTS);   // [SLOTS]
    unsigned char* smap = reinterpret_cast<unsigned char*>(scnt + SLOTS);  // [WARPS][32 * PC_SKM_COUNT_NMAX]
    unsigned short* scand = reinterpret_cast<unsigned short*>(smap + WARPS * 32 * PC_SKM_COUNT_NMAX);   // [PC_SKM_NCAND]
    // flags of the sub-bucket being counted, double-buffered by sub-bucket parity: everybody reads pair [p] after the
    // barrier that ends sub-bucket i while thread 0 already clears pair [p ^ 1] for sub-bucket i + 1 -- no extra barrier
    __shared__ unsigned int s_ovf[2], s_ncand[2], s_next;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    unsigned char* wmap = smap + wid * (32 * PC_SKM_COUNT_NMAX);
    // k-mers that reach keep_min are listed when the count atomic returns keep_min - 2, so the epilogue emits them from
    // the list instead of scanning all SLOTS slots (17 % of this kernel's instructions); keep_min = 1 lists nothing and
    // scans (every distinct k-mer is emitted anyway)
    const bool cand_mode = keep_min >= 2u;

    for (int i = t; i < SLOTS; i += THREADS) { skeys[i] = 0ull; scnt[i] = 0u; }
    // Sub-buckets are handed out by a global counter, not by a fixed stride: minimizer buckets are skewed (one m-mer inside
    // a high-copy repeat family collects the k-mers of every copy), and with static striding the CTA that drew the heavy
    // ones finished long after the others.  A CTA knows the ids of its current, next and next-but-one sub-bucket, so the
    // offsets and the first records of the next one are still requested before the current one is finished.
    if (t == 0) { s_ovf[0] = s_ovf[1] = 0; s_ncand[0] = s_ncand[1] = 0; s_next = atomicAdd(queue, 3u); }
    __syncthreads();

    unsigned long long distinct = 0;
    unsigned int max_probe = 0;
    unsigned int cur_sb = s_next, next_sb = cur_sb + 1, next2_sb = cur_sb + 2;
    if (cur_sb >= n_sub) return;
    // Every warp owns an equal slice of the sub-bucket's records (+-1) and walks it 32 records at a time: with one
    // record per THREAD per pass, a 740-record sub-bucket left half of the CTA's warps without work in its second pass
    // and the epilogue barrier waited for the rest (ncu: barrier was the top stall reason).
    unsigned long long wbeg, wend;                               // this warp's slice of the current sub-bucket
    auto slice = [&](unsigned long long b, unsigned long long e) {
        const unsigned long long per = (e - b + WARPS - 1) / WARPS;
        wbeg = min(e, b + (unsigned long long)wid * per);
        wend = min(e, wbeg + per);
    };
    slice(__ldg(sub_off + cur_sb), __ldg(sub_off + cur_sb + 1));
    unsigned long long cur_i0 = wbeg;
    unsigned long long nbeg = 0, nend = 0;
    if (next_sb < n_sub) { nbeg = __ldg(sub_off + next_sb); nend = __ldg(sub_off + next_sb + 1); }
    SkmRec nrec = make_ulonglong2(0ull, 0ull);
    bool nhave = false;
    {
        const unsigned long long i = cur_i0 + lane;
        nhave = i < wend;
        if (nhave) { const uint4 v = ldnc_u4(recs + i); nrec.x = ((unsigned long long)v.y << 32) | v.x; nrec.y = ((unsigned long long)v.w << 32) | v.z; }
    }
    unsigned int nk = 0;
    unsigned int par = 0;                                        // parity of the sub-bucket this warp is inserting
    for (;;) {
        const SkmRec rec = nrec;
        const bool have = nhave;
        const unsigned int this_sb = cur_sb;
        volatile unsigned int* ovf_flag = &s_ovf[par];
        // (a sub-bucket that has overflowed its table is abandoned at once: the fall-back re-counts all of it, and a heavy
        // minimizer can put millions of records into one sub-bucket -- measured on 4 GPUs: one CTA streaming such a
        // sub-bucket to its end stretched the whole kernel from 6.5 to 11 ms)
        const bool last_batch = cur_i0 + 32 >= wend || *ovf_flag != 0u;
        if (!last_batch) cur_i0 += 32;
        else {
            cur_sb = next_sb; next_sb = next2_sb;              // (next2_sb is refreshed after this sub-bucket's epilogue)
            slice(nbeg, nend);
            cur_i0 = wbeg;
            if (next_sb < n_sub) { nbeg = __ldg(sub_off + next_sb); nend = __ldg(sub_off + next_sb + 1); }
        }
        const bool more = cur_sb < n_sub;
        nhave = false;
        if (more) {
            const unsigned long long i = cur_i0 + lane;
            nhave = i < wend;
            if (nhave) { const uint4 v = ldnc_u4(recs + i); nrec.x = ((unsigned long long)v.y << 32) | v.x; nrec.y = ((unsigned long long)v.w << 32) | v.z; }
        }
        // ---- expand + insert the 32 records of this warp (warp-synchronous: no CTA barrier inside a sub-bucket)
        {
            const int n = have ? min(skm_rec_n(rec.y), PC_SKM_COUNT_NMAX) : 0;
            int incl = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += v;
            }
            const int excl = incl - n;
            const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            for (int j = 0; j < n; ++j) wmap[excl + j] = (unsigned char)lane;
            // the record travels by shuffle with the number of its first k-mer in the (here unused) destination bits
            const unsigned long long rec_lo = (rec.y & ~0x3FFFFFFFull) | (unsigned long long)excl;
            __syncwarp();
            const bool live = *ovf_flag == 0u;
            for (int g0 = 0; g0 < total; g0 += 32 * GRP) {
                unsigned long long skey[GRP], old[GRP]; unsigned int slot[GRP]; bool on[GRP];
#pragma unroll
                for (int u = 0; u < GRP; ++u) {
                    const int g = g0 + u * 32 + lane;
                    on[u] = g < total;
                    const int src = on[u] ? (int)wmap[g] : 0;
                    const unsigned long long hi = __shfl_sync(0xFFFFFFFFu, rec.x, src);
                    const unsigned long long lo = __shfl_sync(0xFFFFFFFFu, rec_lo, src);
                    const uint64_t key = skm_rec_kmer(hi, lo, on[u] ? g - (int)(lo & 0x3FFu) : 0, k);
                    skey[u] = ~key;
                    slot[u] = smem_slot<SLOTS>(key);
                    on[u] = on[u] && live;
                }
#pragma unroll
                for (int u = 0; u < GRP; ++u) old[u] = on[u] ? atomicCAS(&skeys[slot[u]], 0ull, skey[u]) : 0ull;
#pragma unroll
                for (int u = 0; u < GRP; ++u) {
                    if (!on[u]) continue;
                    unsigned long long o = old[u];
                    unsigned int s = slot[u], probes = 0;
                    // (ptxas unrolls this loop up to the probe limit -- 261 ATOMS, 5 600 instructions -- and that is the faster
                    // code: with `#pragma unroll 1` the kernel took 6.7 ms instead of 5.25)
                    for (;;) {
                        if (o == 0ull) { ++nk; break; }
                        if (o == skey[u]) {
                            if (cand_mode) {
                                if (atomicAdd(&scnt[s], 1u) + 2u == keep_min) {           // this occurrence lifts it to keep_min
                                    const unsigned int p = atomicAdd(&s_ncand[par], 1u);
                                    if (p < PC_SKM_NCAND) scand[p] = (unsigned short)s;
                                }
                            } else atomicAdd(&scnt[s], 1u);
                            break;
                        }
                        if (++probes > PC_SKM_PROBE_LIMIT) { *ovf_flag = 1u; break; }
                        s = (s + 1) & (SLOTS - 1);
                        o = atomicCAS(&skeys[s], 0ull, skey[u]);
                    }
                    max_probe = max(max_probe, probes);
                }
            }
            __syncwarp();                                        // the next pass rewrites the map
        }
        if (!last_batch) continue;

        // ---- sub-bucket complete: emit the k-mers that reach keep_min, re-zero the table.  Two CTA barriers: every warp
        // emits and re-zeroes only the slots it owns ([wid * SPW, (wid + 1) * SPW)), so nothing has to be ordered between
        // the emission and the re-zeroing.
        __syncthreads();                                         // A: all inserts of this_sb are done
        const bool ovf = s_ovf[par] != 0u;
        const unsigned int nc = s_ncand[par];
        if (t == 0) {
            if (ovf) ovf_ids[atomicAdd(n_ovf, 1u)] = this_sb;
            s_ovf[par ^ 1u] = 0u; s_ncand[par ^ 1u] = 0u;        // (read by nobody until the next sub-bucket, see above)
            s_next = atomicAdd(queue, 1u);
        }
        if (!ovf) distinct += nk;
        nk = 0;
        const unsigned int lo_slot = (unsigned int)(wid * SPW), hi_slot = lo_slot + SPW;
        if (!ovf && cand_mode && nc <= PC_SKM_NCAND) {
            // the candidate list is short (84 entries per sub-bucket at cfg2): every warp walks it twice and takes its own
            // slots -- first to count them (ONE reservation per warp in the global list: with one per 32 candidates the
            // same-address atomics were what the kernel waited for), then to emit them
            unsigned int mine_n = 0;
            for (unsigned int i0 = 0; i0 < nc; i0 += 32) {
                const unsigned int i = i0 + lane;
                const unsigned int s = i < nc ? (unsigned int)scand[i] : 0xFFFFFFFFu;
                mine_n += __popc(__ballot_sync(0xFFFFFFFFu, s >= lo_slot && s < hi_slot));
            }
            if (mine_n) {
                unsigned long long pos = 0;
                if (lane == 0) pos = atomicAdd(n_list, (unsigned long long)mine_n);
                pos = __shfl_sync(0xFFFFFFFFu, pos, 0);
                for (unsigned int i0 = 0; i0 < nc; i0 += 32) {
                    const unsigned int i = i0 + lane;
                    const unsigned int s = i < nc ? (unsigned int)scand[i] : 0xFFFFFFFFu;
                    const bool mine = s >= lo_slot && s < hi_slot;
                    const unsigned int bal = __ballot_sync(0xFFFFFFFFu, mine);
                    if (mine) {
                        const unsigned long long q = pos + __popc(bal & ((1u << lane) - 1u));
                        if (q < list_cap) { list_keys[q] = ~skeys[s]; list_counts[q] = scnt[s] + 1u; }
                    }
                    pos += __popc(bal);
                }
            }
        } else if (!ovf) {
            // scan of the warp's own slots (keep_min = 1, or more candidates than the list holds)
            uint4 c4[IT4];
            unsigned int emask = 0, n_emit = 0;
            const int s0 = wid * SPW + lane * 4;
#pragma unroll
            for (int i = 0; i < IT4; ++i) {
                const int s = s0 + i * 128;
                c4[i] = *reinterpret_cast<const uint4*>(scnt + s);
                const unsigned int cc[4] = {c4[i].x, c4[i].y, c4[i].z, c4[i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    bool e = cc[j] + 1u >= keep_min;
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
            unsigned long long gb = 0;
            if (lane == 0 && wtot) gb = atomicAdd(n_list, (unsigned long long)wtot);
            gb = __shfl_sync(0xFFFFFFFFu, gb, 0);
            unsigned long long pos = gb + (incl - n_emit);
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
            }
        }
        __syncwarp();                                            // this warp has read what it emits: re-zero its slots
        {
            uint4* cz = reinterpret_cast<uint4*>(scnt + wid * SPW);          // SPW counts = SPW / 4 x 16 bytes
            uint4* kz = reinterpret_cast<uint4*>(skeys + wid * SPW);         // SPW keys   = SPW / 2 x 16 bytes
#pragma unroll
            for (int i = 0; i < SPW / 128; ++i) cz[i * 32 + lane] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int i = 0; i < SPW / 64; ++i) kz[i * 32 + lane] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();                                         // D: table is empty again, s_next is valid
        next2_sb = s_next;
        par ^= 1u;
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

// ------------------------------------------------------------------------------------------------ pass C': expand
// Alternative to expanding inside the count kernel (pc_tune skm_expand 1): the records of every sub-bucket are expanded
// into plain 8-byte canonical k-mers (sub-bucket order, the layout k_count_smem reads), so the count kernel of the k-mer
// path runs unchanged with perfectly balanced threads.  Same warp scheme as k_count_skm (equal record slices per warp,
// k-mers spread evenly over the lanes, direct extraction), but the k-mers leave as fully coalesced 256-byte stores: a
// warp reserves the range of its pass with one global atomic on the sub-bucket's cursor.  No CTA barrier at all.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_rec_expand(const SkmRec* __restrict__ recs, const unsigned long long* __restrict__ sub_off /* records */,
             const unsigned long long* __restrict__ kmer_off, unsigned long long* __restrict__ cursor, unsigned int n_sub,
             int k, uint64_t* __restrict__ out)
{
    constexpr int WARPS = THREADS / 32;
    __shared__ unsigned char smap[WARPS][32 * PC_SKM_COUNT_NMAX];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char* wmap = smap[wid];
    for (unsigned int sb = blockIdx.x; sb < n_sub; sb += gridDim.x) {
        const unsigned long long b = __ldg(sub_off + sb), e = __ldg(sub_off + sb + 1);
        const unsigned long long per = (e - b + WARPS - 1) / WARPS;
        const unsigned long long wbeg = min(e, b + (unsigned long long)wid * per), wend = min(e, wbeg + per);
        const unsigned long long obase = __ldg(kmer_off + sb);
        for (unsigned long long i0 = wbeg; i0 < wend; i0 += 32) {
            const unsigned long long i = i0 + lane;
            SkmRec rec = make_ulonglong2(0ull, 0ull);
            int n = 0;
            if (i < wend) {
                const uint4 v = ldnc_u4(recs + i);
                rec.x = ((unsigned long long)v.y << 32) | v.x; rec.y = ((unsigned long long)v.w << 32) | v.z;
                n = min(skm_rec_n(rec.y), PC_SKM_COUNT_NMAX);
            }
            int incl = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += v;
            }
            const int excl = incl - n;
            const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            for (int j = 0; j < n; ++j) wmap[excl + j] = (unsigned char)lane;
            const unsigned long long rec_lo = (rec.y & ~0x3FFFFFFFull) | (unsigned long long)excl;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&cursor[sb], (unsigned long long)total);
            base = obase + __shfl_sync(0xFFFFFFFFu, base, 0);
            __syncwarp();
            for (int g0 = 0; g0 < total; g0 += 32) {              // warp-uniform trip count: every lane may be a shuffle source
                const int g = g0 + lane;
                const bool on = g < total;
                const int src = on ? (int)wmap[g] : 0;
                const unsigned long long hi = __shfl_sync(0xFFFFFFFFu, rec.x, src);
                const unsigned long long lo = __shfl_sync(0xFFFFFFFFu, rec_lo, src);
                if (on) __stcs(out + base + g, skm_rec_kmer(hi, lo, g - (int)(lo & 0x3FFu), k));
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------ overflow fall-back
// k-mer instances in the listed sub-buckets (sizes the global table of the fall-back): the packed histogram knows them
__global__ void __launch_bounds__(256)
k_hist_kmer_total(const unsigned long long* __restrict__ packed_hist, const unsigned int* __restrict__ ids, unsigned int n_ids,
                  unsigned long long* __restrict__ total)
{
    unsigned long long s = 0;
    for (unsigned int x = blockIdx.x * 256u + threadIdx.x; x < n_ids; x += gridDim.x * 256u) s += packed_hist[ids[x]] >> 32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(total, s);
}

// the listed sub-buckets counted into ONE global open-addressing table (adversarial input or capped sub-bucket counts
// only: results never depend on sizing)
__global__ void __launch_bounds__(256)
k_count_rec_table(const SkmRec* __restrict__ recs, const unsigned long long* __restrict__ sub_off,
                  const unsigned int* __restrict__ ids, unsigned int n_ids, int k, CountTable table, uint64_t capacity,
                  CountStats* __restrict__ stats)
{
    unsigned int new_keys = 0, max_probe = 0;
    for (unsigned int x = blockIdx.x; x < n_ids; x += gridDim.x) {
        const unsigned long long b = sub_off[ids[x]], e = sub_off[ids[x] + 1];
        for (unsigned long long i = b + (unsigned long long)blockIdx.y * 256u + threadIdx.x; i < e; i += 256ull * gridDim.y) {
            const SkmRec rec = recs[i];
            RecRoller r;
            r.init(rec.x, rec.y, k);
            const int n = skm_rec_n(rec.y);
            for (int j = 0; j < n; ++j) {
                const uint64_t key = r.canonical();
                const uint64_t slot = slot_of(hash64(key), capacity);
                count_insert(table, capacity, slot, ~key, ldcg_u64(table.key(slot)), new_keys, max_probe, stats);
                r.step();
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

// ------------------------------------------------------------------------------------------------ K5 compact repeat table
// 8-byte slots for large repeat sets: (32-bit fingerprint << 32) | (column + 1), 0 = empty, load 0.5 -> 16 bytes per repeat
// k-mer instead of 64 (16-byte slots at load 0.25), so the 5.6 M repeat k-mers of the 8-GPU run stay L2 resident (90 MB
// instead of 358 MB).  A fingerprint match is verified against the sorted key array (column = rank), so the look-up is
// exact; only real hits (and one in 2^32 of the other slots met) pay that second read.
__device__ __forceinline__ unsigned int rep_fp32(uint64_t key) { return (unsigned int)((key * 0xC2B2AE3D27D4EB4Full) >> 32); }

__global__ void __launch_bounds__(256)
k_rep8_build(const uint64_t* __restrict__ sorted_keys, int64_t n, unsigned long long* __restrict__ table8, uint64_t cap8)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t key = sorted_keys[i];
    const unsigned long long val = ((unsigned long long)rep_fp32(key) << 32) | (unsigned long long)(i + 1);
    uint64_t s = rep_slot(key, cap8);
    for (;;) {                                               // keys are unique and cap8 >= 2 n: terminates
        if (atomicCAS(&table8[s], 0ull, val) == 0ull) break;
        if (++s == cap8) s = 0;
    }
}

// column of `key` or -1
__device__ __forceinline__ int rep8_find(const unsigned long long* __restrict__ table8, uint64_t cap8,
                                         const uint64_t* __restrict__ sorted_keys, uint64_t key, uint64_t s, unsigned long long v)
{
    const unsigned int fp = rep_fp32(key);
    while (v != 0ull) {
        if ((unsigned int)(v >> 32) == fp) {
            const int col = (int)(unsigned int)v - 1;
            if (__ldg(sorted_keys + col) == key) return col;
        }
        if (++s == cap8) s = 0;
        v = __ldg(table8 + s);
    }
    return -1;
}

// ------------------------------------------------------------------------------------------------ K5 without a k-mer stream
// The super-k-mer count never materialises one 8-byte k-mer per window, so the probe re-derives the canonical k-mers
// from the packed genome (0.375 B per base, L2 resident) with the roller instead of streaming 8 B per window from HBM.
// Same structure as k_probe_stream_filt: persistent CTAs hold the bitmap filter in shared memory, ~2/3 of the windows
// that are no repeat k-mers never leave the SM, the rest probes the 16-byte repeat-table slot in L2.
// SMEMF: the bitmap lives in shared memory (loaded once per CTA); otherwise `bitmap` (may be null) is read through L2.
// Roller of this kernel: the validity of the 32 windows is ONE 32-bit mask (masked-base bits dilated by k - 1, as in the
// super-k-mer extraction) instead of two 64-bit variable shifts per window, and the incoming bases come off the top of a
// shifting register instead of being located in (w0, w1) by index (ncu source page: 11 % of the kernel's instructions).
struct ProbeRoller {
    uint64_t fwd, rc, kmask, t;
    uint32_t vmsb;                                             // bit 31: current window valid (shifted left every step)
    int rcshift;
    __device__ __forceinline__ void init(uint64_t w0, uint64_t w1, uint32_t m0, uint32_t m1, int k) {
        kmask = kmer_mask(k);
        rcshift = 2 * (k - 1);
        fwd = w0 >> (64 - 2 * k);
        rc = revcomp_bits(fwd, k);
        t = k >= 32 ? w1 : ((w0 << (2 * k)) | (w1 >> (64 - 2 * k)));      // bases k, k + 1, ... at the top
        uint64_t r = ((uint64_t)m0 << 32) | (uint64_t)m1;                   // bit (63 - j): base j is masked
        int width = 1;
#pragma unroll
        for (int s = 1; s <= 16; s <<= 1)
            if (2 * s <= k) { r |= r << s; width = 2 * s; }
        r |= r << (k - width);                                              // k - width < width
        vmsb = ~(uint32_t)(r >> 32);
    }
    __device__ __forceinline__ bool valid() const { return (vmsb >> 31) != 0u; }
    __device__ __forceinline__ uint64_t canonical() const { return fwd < rc ? fwd : rc; }
    __device__ __forceinline__ void step() {
        const uint64_t nb = t >> 62;
        t <<= 2;
        vmsb <<= 1;
        fwd = ((fwd << 2) | nb) & kmask;
        rc = (rc >> 2) | ((3ull - nb) << rcshift);
    }
};

template <int BATCH, int THREADS, bool SMEMF, bool TABLE8>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
k_probe_roll_filt(const uint64_t* __restrict__ words, const uint32_t* __restrict__ nmask, int64_t n_words, int k,
                  const int64_t* __restrict__ chunk_word0, int64_t n_chunks, const int32_t* __restrict__ chunk_row,
                  const int64_t* __restrict__ row_hit_base, const RepSlot* __restrict__ rtable, uint64_t rcap,
                  const unsigned short* __restrict__ fp, const unsigned int* __restrict__ bitmap, unsigned int nbits,
                  const unsigned long long* __restrict__ table8, uint64_t cap8, const uint64_t* __restrict__ sorted_keys,
                  int32_t* __restrict__ hits, unsigned int* __restrict__ row_cursor)
{
    extern __shared__ __align__(16) unsigned int sbits[];
    if (SMEMF) {
        const unsigned int n4 = ((nbits + 31) / 32 + 3) / 4;
        const uint4* src = reinterpret_cast<const uint4*>(bitmap);
        uint4* dst = reinterpret_cast<uint4*>(sbits);
        for (unsigned int i = threadIdx.x; i < n4; i += THREADS) dst[i] = __ldg(src + i);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int64_t n_tiles = (n_words + THREADS - 1) / THREADS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t w = tile * THREADS + threadIdx.x;
        const int64_t w_lane0 = w - lane;
        if (w_lane0 >= n_words) continue;                        // whole warp past the end
        WordPair p = load_word_pair(words, nmask, w, n_words);
        int row = -1;
        if (w < n_words && p.m0 != 0xFFFFFFFFu) {
            int64_t c = chunk_of_word(chunk_word0, n_chunks, w, w_lane0);
            row = __ldg(chunk_row + c);
        }
        const bool lane_on = row >= 0;
        const unsigned int on_mask = __ballot_sync(0xFFFFFFFFu, lane_on);
        if (on_mask == 0u) continue;
        const int row0 = __shfl_sync(0xFFFFFFFFu, row, __ffs(on_mask) - 1);
        const bool uniform = __all_sync(0xFFFFFFFFu, !lane_on || row == row0);
        ProbeRoller r;
        if (lane_on) r.init(p.w0, p.w1, p.m0, p.m1, k);
#pragma unroll 1
        for (int b0 = 0; b0 < 32; b0 += BATCH) {
            int cols[BATCH];
            unsigned int hmask = 0;
            if (lane_on) {
                unsigned long long skey[BATCH]; uint64_t slot[BATCH]; uint4 cur[BATCH];
                unsigned int vmask = 0;
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    const uint64_t key = r.canonical();
                    bool maybe = r.valid();
                    if (SMEMF) {
                        const unsigned int fb = filt_bit(key, nbits);
                        maybe = maybe && ((sbits[fb >> 5] >> (fb & 31)) & 1u);
                    }
                    vmask |= (unsigned int)maybe << j;
                    skey[j] = ~key;
                    slot[j] = rep_slot(key, TABLE8 ? cap8 : rcap);
                    r.step();                                    // (the step after window 31 shifts in zeros: never used)
                }
                if (!SMEMF && bitmap != nullptr) {               // L2-resident bitmap (repeat sets too large for shared memory)
                    unsigned int wbits[BATCH], fb[BATCH];
#pragma unroll
                    for (int j = 0; j < BATCH; ++j) {
                        fb[j] = gbit_of(~skey[j], nbits);
                        wbits[j] = ((vmask >> j) & 1u) ? __ldg(bitmap + (fb[j] >> 5)) : 0u;
                    }
#pragma unroll
                    for (int j = 0; j < BATCH; ++j)
                        if (!((wbits[j] >> (fb[j] & 31)) & 1u)) vmask &= ~(1u << j);
                }
                if (TABLE8) {                                    // compact 8-byte slots, verified against the key array
                    unsigned long long v8[BATCH];
#pragma unroll
                    for (int j = 0; j < BATCH; ++j) v8[j] = ((vmask >> j) & 1u) ? __ldg(table8 + slot[j]) : 0ull;
#pragma unroll
                    for (int j = 0; j < BATCH; ++j) {
                        cols[j] = 0;
                        if (v8[j] == 0ull) continue;
                        const int col = rep8_find(table8, cap8, sorted_keys, ~skey[j], slot[j], v8[j]);
                        if (col >= 0) { cols[j] = col; hmask |= 1u << j; }
                    }
                    vmask = 0;
                } else if (fp != nullptr) {                      // fingerprint-first walk (tables far beyond L2)
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
                    vmask = 0;
                }
#pragma unroll
                for (int j = 0; j < BATCH; ++j)
                    cur[j] = ((vmask >> j) & 1u) ? __ldg(reinterpret_cast<const uint4*>(rtable + slot[j])) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    if (fp == nullptr && !TABLE8) cols[j] = 0;
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
            const int nh = __popc(hmask);
            const unsigned int any = __ballot_sync(0xFFFFFFFFu, nh > 0);
            if (any == 0u) continue;
            int32_t* dst = nullptr;
            if (uniform) {
                int incl = nh;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (lane >= o) incl += v;
                }
                const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
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

}  // namespace pc
