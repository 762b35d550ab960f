// K6 without a sort: rows -> CSR through a shared-memory column bitmap (replaces k_row_sort + k_indptr + k_row_emit,
// which replaced bedtools sort/merge + Counter + dok_matrix.tocsc of polycracker.py:255-281, 299-367).
//
// A row's hits are column ids (ranks in the sorted repeat set), unordered, with multiplicity.  What CSR wants is the
// DISTINCT columns in ascending order with their multiplicities.  A bitmap over the columns gives exactly that order for
// free: set one bit per hit, and the rank of a column among the row's distinct columns is the number of set bits below
// it.  One persistent CTA (1024 threads, the whole shared memory of its SM) handles one row at a time:
//   sweep A   every hit sets its bit (shared atomicOr); population count of the bitmap = the row's entry count
//   look-back the row publishes its entry count and takes the sum of all earlier rows (rows are handed out in order by a
//             global ticket, so the single-pass chained scan cannot dead-lock): the row's CSR offset, no second kernel
//   sweep B   every hit ranks its column (set bits below it: per-chunk and per-word prefixes kept next to the bitmap, three
//             shared loads and one POPC) and adds 1 to counts[rank] (shared memory; ranks beyond the shared array -- a row
//             with more distinct columns than fit -- count in the row's own slice of `data` in global memory)
//   emit      the bitmap is walked one chunk (32 words) per warp: columns and df += 1 leave in ascending order and the
//             words are cleared on the way; the counts follow as one coalesced copy.
// Repeat sets larger than one bitmap are handled in power-of-two column ranges: the row's hits are first grouped by range
// (a counting sort into the second hit buffer, which exists only then), and every range runs both sweeps over its own
// hits.  The hits of one row are <= 300 KB and stay in L2 between the sweeps.  For a single range the row's hits are
// read straight from the probe's hit buffer: no second buffer, none of the sort's global ping-pong passes.
#pragma once
#include "kernels.cuh"

namespace pc {

#define PC_CSR_THREADS 1024
#define PC_CSR_MAXR 64                                         // column ranges per row the kernel supports
#define PC_CSR_FLAG_AGG (1ull << 62)
#define PC_CSR_FLAG_INCL (2ull << 62)
#define PC_CSR_VAL_MASK ((1ull << 62) - 1ull)

// sum of v over the CTA (every thread gets it); part = 32 words of shared memory, two barriers
__device__ __forceinline__ unsigned int csr_cta_sum(unsigned int v, unsigned int* part) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (lane == 0) part[wid] = v;
    __syncthreads();
    unsigned int s = part[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    __syncthreads();
    return s;
}

// exclusive prefix of v over the CTA in thread order; part as above
__device__ __forceinline__ unsigned int csr_cta_excl(unsigned int v, unsigned int* part) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) part[wid] = incl;
    __syncthreads();
    unsigned int w = part[lane], winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int u = __shfl_up_sync(0xFFFFFFFFu, winc, o);
        if (lane >= o) winc += u;
    }
    const unsigned int before = __shfl_sync(0xFFFFFFFFu, winc - w, wid);
    __syncthreads();
    return before + incl - v;
}

__global__ void __launch_bounds__(256)
k_sum_u32(const unsigned int* __restrict__ v, int64_t n, unsigned long long* __restrict__ out)
{
    unsigned long long s = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// The hits of a row are walked with CSR_U loads in flight per thread: with one load per iteration (the shared atomic
// that follows keeps the compiler from hoisting the next one) a 45 k-hit row cost 44 dependent DRAM round trips per
// sweep -- measured 41 us per row, 1.84 ms for the 6 621 rows of cfg2.
#define PC_CSR_U 8
#define PC_CSR_FOR_HITS(PTR, N, C0, LIMIT, ...)                                                                \
    for (unsigned int i0_ = t; i0_ < (N); i0_ += T * PC_CSR_U) {                                               \
        unsigned int cc_[PC_CSR_U];                                                                            \
        _Pragma("unroll")                                                                                      \
        for (int u_ = 0; u_ < PC_CSR_U; ++u_) {                                                                \
            const unsigned int i_ = i0_ + u_ * T;                                                              \
            cc_[u_] = i_ < (N) ? (unsigned int)__ldcg((PTR) + i_) - (C0) : 0xFFFFFFFFu;                         \
        }                                                                                                      \
        _Pragma("unroll")                                                                                      \
        for (int u_ = 0; u_ < PC_CSR_U; ++u_) {                                                                \
            const unsigned int c = cc_[u_];                                                                    \
            if (c < (LIMIT)) { __VA_ARGS__ }                                                                   \
        }                                                                                                      \
    }

__global__ void __launch_bounds__(PC_CSR_THREADS, 1)
k_row_bitmap_csr(const int32_t* __restrict__ hits, const int64_t* __restrict__ row_hit_base,
                 const unsigned int* __restrict__ row_cursor, int n_rows, unsigned int n_cols,
                 unsigned int range_bits /* multiple of 1024, <= 1024 * PC_CSR_THREADS */, unsigned int ncap /* shared count slots */,
                 unsigned long long* row_state /* [n_rows], zeroed */, unsigned int* queue /* zeroed */,
                 int64_t* __restrict__ indptr, int32_t* __restrict__ indices, float* data, int* __restrict__ df,
                 int32_t* scratch /* several ranges: a second hit buffer, same row regions; range_bits = 1 << range_shift */,
                 unsigned int range_shift, unsigned int cnt_alloc /* count slots laid out: >= ncap and >= 32 * ranges */)
{
    constexpr int T = PC_CSR_THREADS, WARPS = T / 32;
    // Bitmap layout: the range is cut into chunks of 32 words (1024 columns); word j of chunk g lives at j * NC + g with
    // NC odd.  A THREAD that walks the 32 words of its own chunk (prefix pass) and a WARP whose lanes hold the 32 words
    // of one chunk (emission) both touch 32 different banks.  rel[] (same indexing) holds the number of set bits of the
    // chunk below each word, cpre[] the number of set bits of the range below each chunk: the rank of a column is
    // cpre[chunk] + rel[word] + popc(word & below) -- three shared loads, no warp scan anywhere.
    const unsigned int nchunks = range_bits / 1024, NC = nchunks | 1u;
    extern __shared__ __align__(16) unsigned int csr_smem[];
    unsigned int* bits = csr_smem;                               // [32 * NC]
    unsigned int* cpre = bits + 32 * NC;                         // [nchunks + 1]
    unsigned int* cnt = cpre + nchunks + 1;                      // [ncap]
    unsigned short* rel = reinterpret_cast<unsigned short*>(cnt + cnt_alloc);   // [32 * NC]
    __shared__ unsigned int part[32];
    __shared__ unsigned int s_nnzp[PC_CSR_MAXR];
    __shared__ unsigned int s_rstart[PC_CSR_MAXR + 1];           // several ranges: where each range's hits start in the scratch row
    __shared__ int s_row;
    __shared__ unsigned long long s_base;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const unsigned int nr = (n_cols + range_bits - 1) / range_bits;
    int* idata = reinterpret_cast<int*>(data);
#define PC_CSR_PHYS(c) ((((c) >> 5) & 31u) * NC + ((c) >> 10))

    for (unsigned int w = t; w < 32 * NC; w += T) bits[w] = 0u;
    for (;;) {
        if (t == 0) s_row = (int)atomicAdd(queue, 1u);
        __syncthreads();
        const int row = s_row;
        if (row >= n_rows) break;
        const unsigned int n = __ldg(row_cursor + row);
        const int32_t* a = hits + __ldg(row_hit_base + row);

        // ---- several ranges: group the row's hits by range first (counting sort into the scratch row; per-warp counters
        // in the not yet used count array), so that every range sweeps only its own hits.  Without it every sweep walked
        // all hits of the row: 6.8 M columns (9 ranges) took 8.65 ms against 4.78 ms for the per-row radix sort.
        if (nr > 1 && n) {
            unsigned int* wc = cnt;                              // [WARPS][nr]
            for (unsigned int i = t; i < WARPS * nr; i += T) wc[i] = 0u;
            __syncthreads();
            PC_CSR_FOR_HITS(a, n, 0u, 0xFFFFFFFFu, atomicAdd(&wc[wid * nr + (c >> range_shift)], 1u);)
            __syncthreads();
            if ((unsigned int)t < nr) {                          // hits of range t, then the per-warp cursors
                unsigned int sum = 0;
                for (int w = 0; w < WARPS; ++w) { const unsigned int v = wc[w * nr + t]; wc[w * nr + t] = sum; sum += v; }
                s_nnzp[t] = sum;                                 // (borrowed: overwritten with the entry counts below)
            }
            __syncthreads();
            if (t == 0) {
                unsigned int run = 0;
                for (unsigned int r = 0; r < nr; ++r) { s_rstart[r] = run; run += s_nnzp[r]; }
                s_rstart[nr] = run;
            }
            __syncthreads();
            int32_t* sr = scratch + (a - hits);
            PC_CSR_FOR_HITS(a, n, 0u, 0xFFFFFFFFu,
                const unsigned int r = c >> range_shift;
                sr[s_rstart[r] + atomicAdd(&wc[wid * nr + r], 1u)] = (int32_t)c;
            )
            __syncthreads();
            a = sr;
        }

        // ---- sweep A: distinct columns per range (the bitmap is all-zero here)
        unsigned int total = 0;
        for (unsigned int p = 0; p < nr; ++p) {
            unsigned int np = 0;
            const int32_t* ap = nr > 1 ? a + s_rstart[p] : a;
            const unsigned int npn = nr > 1 ? (n ? s_rstart[p + 1] - s_rstart[p] : 0u) : n;
            if (npn) {
                const unsigned int c0 = p * range_bits;
                PC_CSR_FOR_HITS(ap, npn, c0, range_bits, atomicOr(&bits[PC_CSR_PHYS(c)], 1u << (c & 31u));)
                __syncthreads();
                unsigned int s = 0;
                if ((unsigned int)t < nchunks) {
#pragma unroll 8
                    for (unsigned int j = 0; j < 32; ++j) {
                        const unsigned int v = bits[j * NC + t];
                        rel[j * NC + t] = (unsigned short)s;
                        s += __popc(v);
                        if (nr > 1 && v) bits[j * NC + t] = 0u;   // several ranges: sweep B sets the bits again
                    }
                }
                const unsigned int ex = csr_cta_excl(s, part);
                if ((unsigned int)t < nchunks) cpre[t] = ex;
                if (t == T - 1) cpre[nchunks] = ex + s;          // (nchunks <= T; thread T - 1 holds the grand total)
                __syncthreads();
                np = cpre[nchunks];
            }
            __syncthreads();                                     // (s_nnzp[p] held the hit count of range p until here)
            if (t == 0) s_nnzp[p] = np;
            total += np;
        }

        // ---- chained scan over the rows (decoupled look-back, warp 0)
        if (wid == 0) {
            unsigned long long excl = 0;
            if (row == 0) {
                if (lane == 0) atomicExch(&row_state[0], PC_CSR_FLAG_INCL | (unsigned long long)total);
            } else {
                if (lane == 0) atomicExch(&row_state[row], PC_CSR_FLAG_AGG | (unsigned long long)total);
                int j = row - 1;
                for (;;) {
                    const int idx = j - lane;
                    unsigned long long v = PC_CSR_FLAG_INCL;     // rows before the first one: inclusive prefix 0
                    if (idx >= 0) {
                        volatile unsigned long long* q = row_state + idx;
                        do { v = *q; } while ((v >> 62) == 0ull);
                    }
                    const unsigned int incl_mask = __ballot_sync(0xFFFFFFFFu, (v >> 62) == 2ull);
                    const int first = incl_mask ? (__ffs(incl_mask) - 1) : 31;     // nearest row with an inclusive prefix
                    unsigned long long s = lane <= first ? (v & PC_CSR_VAL_MASK) : 0ull;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
                    excl += s;
                    if (incl_mask) break;
                    j -= 32;
                }
                if (lane == 0) atomicExch(&row_state[row], PC_CSR_FLAG_INCL | (excl + (unsigned long long)total));
            }
            if (lane == 0) {
                s_base = excl;
                indptr[row + 1] = (int64_t)(excl + total);
                if (row == 0) indptr[0] = 0;
            }
        }
        __syncthreads();
        unsigned long long run = s_base;

        // ---- sweep B: counts per distinct column, emission in column order
        for (unsigned int p = 0; p < nr; ++p) {
            const unsigned int np = s_nnzp[p];
            if (np == 0) continue;                               // (uniform; the bitmap is still all-zero)
            const unsigned int c0 = p * range_bits;
            const int32_t* ap = nr > 1 ? a + s_rstart[p] : a;
            const unsigned int npn = nr > 1 ? s_rstart[p + 1] - s_rstart[p] : n;
            if (nr > 1) {
                PC_CSR_FOR_HITS(ap, npn, c0, range_bits, atomicOr(&bits[PC_CSR_PHYS(c)], 1u << (c & 31u));)
                __syncthreads();
                unsigned int s = 0;
                if ((unsigned int)t < nchunks) {
#pragma unroll 8
                    for (unsigned int j = 0; j < 32; ++j) {
                        rel[j * NC + t] = (unsigned short)s;
                        s += __popc(bits[j * NC + t]);
                    }
                }
                const unsigned int ex = csr_cta_excl(s, part);
                if ((unsigned int)t < nchunks) cpre[t] = ex;
            }
            const unsigned int m = min(np, ncap);
            for (unsigned int i = t; i < m; i += T) cnt[i] = 0u;
            for (unsigned int i = ncap + t; i < np; i += T) idata[run + i] = 0;
            __syncthreads();
            PC_CSR_FOR_HITS(ap, npn, c0, range_bits,
                const unsigned int ph = PC_CSR_PHYS(c);
                const unsigned int pos = cpre[c >> 10] + rel[ph] + __popc(bits[ph] & ((1u << (c & 31u)) - 1u));
                if (pos < ncap) atomicAdd(&cnt[pos], 1u);
                else atomicAdd(idata + run + pos, 1);
            )
            __syncthreads();
            // columns + df: one warp per chunk, one word per lane (bank-conflict free, see above); the bitmap is cleared
            for (unsigned int g = wid; g < nchunks; g += WARPS) {
                const unsigned int ph = lane * NC + g;
                unsigned int word = bits[ph];
                if (word == 0u) continue;
                bits[ph] = 0u;
                int32_t* out = indices + run + cpre[g] + rel[ph];
                const unsigned int colbase = c0 + g * 1024u + lane * 32u;
                do {
                    const unsigned int col = colbase + (__ffs(word) - 1);
                    word &= word - 1u;
                    *out++ = (int32_t)col;
                    atomicAdd(&df[col], 1);
                } while (word);
            }
            // counts: one entry per thread, coalesced
            for (unsigned int i = t; i < m; i += T) data[run + i] = (float)cnt[i];
            for (unsigned int i = ncap + t; i < np; i += T) data[run + i] = (float)__ldcg(idata + run + i);
            __syncthreads();
            run += np;
        }
    }
#undef PC_CSR_PHYS
}

}  // namespace pc
