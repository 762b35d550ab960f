// N3, second half: the Gram matrix K += P P^T of a dense float32 column panel P (rows x panel columns, row-major) on the
// 5th-generation tensor cores -- the one dense contraction on this path (sklearn's linear_kernel / cosine_similarity
// behind KernelPCA, polycracker.py:387, 398-399).  Hand-written for sm_100a:
//   * operands: TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) brings 128 rows x 32 float32 (= one 128-byte swizzle row per
//     matrix row) of the A tile (rows of block i) and of the B tile (rows of block j) into shared memory, three stages,
//     mbarrier full / empty pairs.  Both operands are K-major views of the SAME panel: one tensor map.
//   * MMA: one elected thread issues tcgen05.mma.cta_group::1.kind::tf32, M = 128, N = 128, K = 8 per instruction (four
//     per stage, the shared-memory descriptor advanced by 32 bytes inside the swizzle atom), accumulator = 128 lanes x
//     128 columns of TMEM; tcgen05.commit releases the stage / announces the finished accumulator.
//   * epilogue: four warps read their 32 TMEM lanes with tcgen05.ld.32x32b.x32 and add the tile to K in global memory
//     (one CTA owns a tile per launch: plain read-modify-write).
// Only tiles with j >= i are computed (K is symmetric); k_gram_mirror copies the upper triangle down at the end.
// Warp roles: 0 = TMA producer, 1 = TMEM allocation + MMA issue, 2..5 = epilogue.  Two CTAs fit one SM (97 KB of shared
// memory, 128 TMEM columns each), so one CTA's epilogue overlaps the other's main loop.
// TF32 keeps 10 mantissa bits: the panel is rounded to nearest TF32 first (by the panel kernel, or k_round_tf32) so the products are unbiased;
// K is within 1e-3 relative of the float64 result (tests/test_gpu.py::test_n3_gram_matrix).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pc {
namespace gram {

constexpr int BM = 128, BN = 128, BK = 32;                     // BK float32 = 128 bytes = one SWIZZLE_128B row
constexpr int STAGES = 3;
constexpr int UMMA_K = 8;                                      // kind::tf32: 32 bytes of K per instruction
constexpr uint32_t TILE_BYTES = BM * BK * 4;                   // 16 KB per operand tile
constexpr uint32_t TMEM_COLS = 128;                            // 128 x 128 float32 accumulator
constexpr int THREADS = 192;
constexpr size_t SMEM_BYTES = 1024 /* alignment slack */ + (size_t)STAGES * 2 * TILE_BYTES + 256 /* barriers */;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// bounded spin: a transfer that never arrives (a bad tensor map) must trap, not hang the device
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    for (unsigned int spins = 0;; ++spins) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
        if (done) return;
        if (spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c_inner, int c_outer) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
                 : "memory");
}
// shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);                  // bits  0..13  start address >> 4
    d |= (uint64_t)1 << 16;                                    // bits 16..29  leading byte offset >> 4 (unused with a swizzle)
    d |= (uint64_t)(1024u >> 4) << 32;                         // bits 32..45  stride byte offset >> 4
    d |= (uint64_t)1 << 46;                                    // bits 46..47  descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                                    // bits 61..63  SWIZZLE_128B
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// round-to-nearest TF32 in place (the tensor core would truncate the low 13 mantissa bits)
__global__ void __launch_bounds__(256)
k_round_tf32(float* __restrict__ p, int64_t n_rows, int ncols, int64_t ld)
{
    const int64_t total = n_rows * (int64_t)ncols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / ncols, c = i - r * ncols;
        float* q = p + r * ld + c;
        uint32_t o;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(o) : "f"(*q));
        *q = __uint_as_float(o);
    }
}

// tile t of the upper triangle (row-major enumeration of the pairs ti <= tj over nt row blocks)
__device__ __forceinline__ void tile_of(int t, int nt, int& ti, int& tj) {
    int i = 0, rowlen = nt;
    while (t >= rowlen) { t -= rowlen; ++i; --rowlen; }
    ti = i; tj = i + t;
}

__global__ void __launch_bounds__(THREADS)
k_gram_tf32(const __grid_constant__ CUtensorMap tm, float* __restrict__ K, int n_rows, int64_t ldk, int n_kblocks, int nt)
{
    extern __shared__ uint8_t gram_smem_raw[];
    uint8_t* smem = gram_smem_raw + ((1024u - (smem_u32(gram_smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B: 1024-byte aligned tiles
    uint8_t* tiles = smem;                                      // [STAGES][A 16 KB | B 16 KB], 1024-byte aligned
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * 2 * TILE_BYTES);   // [STAGES]
    uint64_t* empty = full + STAGES;                            // [STAGES]
    uint64_t* accum = empty + STAGES;                           // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int ti, tj;
    tile_of((int)blockIdx.x, nt, ti, tj);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm)) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                            // one warp allocates (and later frees) the TMEM columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                        // ---- TMA producer
            for (int kb = 0; kb < n_kblocks; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);                  // (a fresh barrier passes the wait on parity 1)
                mbar_expect_tx(&full[s], 2 * TILE_BYTES);
                uint8_t* a = tiles + (size_t)s * 2 * TILE_BYTES;
                tma_load_2d(a, &tm, &full[s], kb * BK, ti * BM);
                tma_load_2d(a + TILE_BYTES, &tm, &full[s], kb * BK, tj * BN);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                        // ---- MMA issue (one thread)
            for (int kb = 0; kb < n_kblocks; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                mbar_wait(&full[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(tiles + (size_t)s * 2 * TILE_BYTES);
                const uint64_t adesc = smem_desc(a_addr), bdesc = smem_desc(a_addr + TILE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)           // +32 bytes of K per instruction: +2 in the address field
                    umma_tf32(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), (uint32_t)((kb | k) != 0));
                umma_commit(&empty[s]);                         // the stage is free once these MMAs have read it
            }
            umma_commit(accum);                                 // accumulator complete
        }
    } else {                                                    // ---- epilogue: warps 2..5 own TMEM lanes 32 * (warp & 3) ...
        mbar_wait(accum, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int g = warp & 3;
        const int row = ti * BM + g * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(c * 32), v);
            const int col0 = tj * BN + c * 32;
            if (row < n_rows) {
                float* out = K + (int64_t)row * ldk + col0;
                if (col0 + 32 <= n_rows && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float4 o = *reinterpret_cast<float4*>(out + 4 * q);
                        o.x += __uint_as_float(v[4 * q]); o.y += __uint_as_float(v[4 * q + 1]);
                        o.z += __uint_as_float(v[4 * q + 2]); o.w += __uint_as_float(v[4 * q + 3]);
                        *reinterpret_cast<float4*>(out + 4 * q) = o;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 32; ++q)
                        if (col0 + q < n_rows) out[q] += __uint_as_float(v[q]);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ persistent variant
// The kernel above pulls 32 KB of operands from L2 for every 128 x 128 x 32 block of products: 32 flop per byte, and ncu shows
// the tensor pipe only 22 % busy while the CTAs wait for their tiles.  This variant raises the reuse and overlaps the
// epilogue: one persistent CTA per SM walks 128 x 256 tiles (one tcgen05.mma of N = 256 per K step: 48 KB per block of
// products, 43.7 flop per byte), four stages, and TWO accumulators of 256 TMEM columns -- the epilogue warps drain one
// while the MMA thread fills the other (tmem_full / tmem_empty barriers).  Tiles with any column on or above the diagonal
// are computed: column block tJ >= ti / 2.
namespace v2 {
constexpr int BN2 = 256;
constexpr int STAGES2 = 4;
constexpr uint32_t A_BYTES = BM * BK * 4, B_BYTES = BN2 * BK * 4;          // 16 KB + 32 KB per stage
constexpr uint32_t TMEM_COLS2 = 512;                                          // two 128 x 256 float32 accumulators
constexpr size_t SMEM_BYTES2 = 1024 + (size_t)STAGES2 * (A_BYTES + B_BYTES) + 256;
constexpr uint32_t IDESC2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN2 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32_n256(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC2), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// tile t of the trapezoid: row block ti (128 rows), column block tJ (256 columns), tJ >= ti / 2
__device__ __forceinline__ void tile_of2(int t, int ntn, int& ti, int& tj) {
    int i = 0, rowlen = ntn;
    while (t >= rowlen) { t -= rowlen; ++i; rowlen = ntn - (i >> 1); }
    ti = i; tj = (i >> 1) + t;
}

__global__ void __launch_bounds__(THREADS, 1)
k_gram_tf32_p(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, float* __restrict__ K, int n_rows,
              int64_t ldk, int n_kblocks, int ntn, int n_tiles)
{
    extern __shared__ uint8_t gram_smem_raw[];
    uint8_t* smem = gram_smem_raw + ((1024u - (smem_u32(gram_smem_raw) & 1023u)) & 1023u);
    uint8_t* tiles = smem;                                      // [STAGES2][A 16 KB | B 32 KB]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES2 * (A_BYTES + B_BYTES));
    uint64_t* empty = full + STAGES2;
    uint64_t* tfull = empty + STAGES2;                          // [2] accumulator complete
    uint64_t* tempty = tfull + 2;                               // [2] accumulator drained (4 epilogue warps arrive)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_b)) : "memory");
        for (int s = 0; s < STAGES2; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS2) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                        // ---- TMA producer
            uint32_t it = 0;                                    // k-blocks issued so far (stage = it % STAGES2)
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                int ti, tj;
                tile_of2(tile, ntn, ti, tj);
                for (int kb = 0; kb < n_kblocks; ++kb, ++it) {
                    const uint32_t s = it % STAGES2, ph = (it / STAGES2) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
                    uint8_t* a = tiles + (size_t)s * (A_BYTES + B_BYTES);
                    tma_load_2d(a, &tm_a, &full[s], kb * BK, ti * BM);
                    tma_load_2d(a + A_BYTES, &tm_b, &full[s], kb * BK, tj * BN2);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                        // ---- MMA issue
            uint32_t it = 0, nt = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++nt) {
                const uint32_t acc = nt & 1u, aph = (nt >> 1) & 1u;
                mbar_wait(&tempty[acc], aph ^ 1u);              // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + acc * (uint32_t)BN2;
                for (int kb = 0; kb < n_kblocks; ++kb, ++it) {
                    const uint32_t s = it % STAGES2, ph = (it / STAGES2) & 1u;
                    mbar_wait(&full[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = smem_u32(tiles + (size_t)s * (A_BYTES + B_BYTES));
                    const uint64_t adesc = smem_desc(a_addr), bdesc = smem_desc(a_addr + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_tf32_n256(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), (uint32_t)((kb | k) != 0));
                    umma_commit(&empty[s]);
                }
                umma_commit(&tfull[acc]);
            }
        }
    } else {                                                    // ---- epilogue warps 2..5
        const int g = warp & 3;
        uint32_t nt = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++nt) {
            int ti, tj;
            tile_of2(tile, ntn, ti, tj);
            const uint32_t acc = nt & 1u, aph = (nt >> 1) & 1u;
            mbar_wait(&tfull[acc], aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = ti * BM + g * 32 + lane;
#pragma unroll 1
            for (int c = 0; c < BN2 / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(g * 32) << 16) + acc * (uint32_t)BN2 + (uint32_t)(c * 32), v);
                const int col0 = tj * BN2 + c * 32;
                if (row < n_rows && col0 < n_rows) {
                    float* out = K + (int64_t)row * ldk + col0;
                    if (col0 + 32 <= n_rows && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float4 o = *reinterpret_cast<float4*>(out + 4 * q);
                            o.x += __uint_as_float(v[4 * q]); o.y += __uint_as_float(v[4 * q + 1]);
                            o.z += __uint_as_float(v[4 * q + 2]); o.w += __uint_as_float(v[4 * q + 3]);
                            *reinterpret_cast<float4*>(out + 4 * q) = o;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 32; ++q)
                            if (col0 + q < n_rows) out[q] += __uint_as_float(v[q]);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);           // this warp's 32 lanes of the accumulator are free again
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS2) : "memory");
    }
}

// K[j][i] = K[i][j] for every j > i (the persistent kernel computes every entry on or above the diagonal, and a few below
// it inside the tiles that straddle it: those are overwritten with their mirror image)
__global__ void __launch_bounds__(256)
k_gram_mirror_all(float* __restrict__ K, int n, int64_t ldk)
{
    const int64_t total = (int64_t)n * n;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / n), j = (int)(idx - (int64_t)i * n);
        if (j > i) K[(int64_t)j * ldk + i] = K[(int64_t)i * ldk + j];
    }
}
}  // namespace v2

// K[j][i] = K[i][j] for j > i (the kernel above fills the upper triangle of every tile pair ti <= tj; inside a
// diagonal tile both halves are computed)
__global__ void __launch_bounds__(256)
k_gram_mirror(float* __restrict__ K, int n, int64_t ldk)
{
    const int64_t total = (int64_t)n * n;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / n), j = (int)(idx - (int64_t)i * n);
        if (j / BN > i / BM) K[(int64_t)j * ldk + i] = K[(int64_t)i * ldk + j];
    }
}

}  // namespace gram
}  // namespace pc
