// C ABI of the B200 hot path (include/polycracker_b200.h): host orchestration of the kernels in kernels.cuh.
// No CPU fallback lives here: every compute entry point needs a CUDA device and says so when there is none.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/polycracker_b200.h"
#include "kernels.cuh"
#include "kernels_smem.cuh"
#include "kernels_skm.cuh"
#include "kernels_csr.cuh"
#include "gram_tc.cuh"

namespace {

// PC_TRACE=1: host wall-clock trace points on stderr (development aid)
static bool trace_on() { static int v = -1; if (v < 0) { const char* e = getenv("PC_TRACE"); v = (e && *e == '1') ? 1 : 0; } return v == 1; }
static double now_ms() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
#define TRACE(label) do { if (trace_on()) fprintf(stderr, "[pc_trace] %-28s %.3f ms\n", label, now_ms()); } while (0)

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}

#define CU(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(PC_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
    } while (0)

enum Stage { ST_H2D = 0, ST_PACK, ST_TABLE_INIT, ST_COUNT, ST_SELECT, ST_SORT, ST_REPTABLE, ST_PROBE,
             ST_ROWSORT, ST_EMIT, ST_TFIDF, ST_D2H, ST_EXPORT, ST_MERGE, ST_PART_HIST, ST_PART_SCATTER,
             ST_SUB_DICT, ST_SUB_COMPARE, ST_SUB_MAPBACK, ST_PART2, ST_COUNT_FALLBACK, ST_GRAM, ST_EXPAND, ST_N };
static_assert(ST_N <= PC_N_STAGES, "stage table too small");
static_assert(pc::kMaxGroups == PC_MAX_GROUPS && pc::kMaxSubK == PC_MAX_SUB_K, "header and kernels disagree");

// Memory pressure: buffers are cached across calls (a step allocates nothing in steady state), so a later pass with a
// different shape -- another k on a 4.5 Gbp genome -- can find HBM full of buffers nobody needs any more.  While pc_count
// runs, a failed cudaMalloc frees what is stale by construction (the previous pass's repeat set and matrix, the other
// count path's partition buffers) and tries once more.  Defined after pc_ctx.
bool relieve_pressure();

template <typename T>
struct DevBuf {
    T* p = nullptr; size_t n = 0;
    int alloc(size_t count) {
        if (count <= n && p) return PC_OK;
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e != cudaSuccess) {
            cudaGetLastError();
            if (relieve_pressure()) e = cudaMalloc(&p, count * sizeof(T));
        }
        if (e != cudaSuccess) { cudaGetLastError(); p = nullptr; n = 0; return fail(PC_ENOMEM, "cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e)); }
        n = count;
        return PC_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

}  // namespace

struct pc_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    uint64_t zeroed_slots = 0;                  // slots of `table` zeroed by the pending side-stream memset
    int prezero = 0;                            // overlap the table memset with extract+scatter (measured: no gain)
    int probe_fp = -1;                          // fingerprint-first probing: -1 auto (table > 256 MB), 0 off, 1 on
    cudaEvent_t ev[ST_N][2];
    bool ev_used[ST_N];
    int64_t launches = 0;
    // read-back mailbox: page-locked host memory the device writes small results into (see read_small)
    uint8_t* mailbox = nullptr; size_t mailbox_cap = 0, mailbox_used = 0;
    struct PendingRead { void* dst; size_t off, bytes; };
    std::vector<PendingRead> pending_reads;

    // sequences
    DevBuf<uint8_t> ascii; int64_t ascii_padded = 0;
    std::vector<int64_t> h_word0, h_begin, h_len;
    DevBuf<int64_t> d_word0, d_begin, d_len;
    int64_t n_chunks = 0, n_words = 0;           // n_words = data words (a guard word follows)
    DevBuf<uint64_t> words; DevBuf<uint32_t> nmask;
    bool loaded = false;

    // count table
    int k = 0;
    DevBuf<unsigned long long> table_keys; DevBuf<unsigned int> table_counts; uint64_t capacity = 0;
    pc::CountTable table() const { return pc::CountTable{table_keys.p, table_counts.p}; }
    DevBuf<pc::CountStats> d_stats; pc::CountStats h_stats{};
    bool counted = false;

    // selection
    DevBuf<uint64_t> sel_a, sel_b; uint64_t* sel = nullptr; int64_t n_sel = 0;
    DevBuf<unsigned int> rs_hist;
    DevBuf<unsigned long long> d_counter;
    DevBuf<pc::RepSlot> rtable; DevBuf<unsigned short> rfp; uint64_t rcap = 0;
    DevBuf<unsigned long long> rtable8; uint64_t rcap8 = 0;     // compact 8-byte slots (large repeat sets, roller probe)
    int probe_table8 = -1;                      // tunable: -1 auto (repeat sets beyond the shared-memory filter, no k-mer stream), 0 off, 1 on
    bool selected = false;

    // matrix
    DevBuf<int32_t> d_chunk_row; DevBuf<int64_t> d_row_base; DevBuf<unsigned int> d_row_cursor;
    DevBuf<int32_t> hits_a, hits_b; DevBuf<int64_t> d_row_nnz, d_indptr;
    DevBuf<int32_t> d_indices; DevBuf<float> d_data, d_idf, d_tfidf; DevBuf<int> d_df, d_df_global;
    DevBuf<double> d_row_norm; DevBuf<unsigned short> d_data16; DevBuf<int64_t> d_esc_pos; DevBuf<float> d_esc_val;
    int64_t n_esc = -1;                          // escapes of the last pc_csr_get16 (-1: none taken)
    DevBuf<unsigned char> d_dcol8, d_cnt8; DevBuf<int64_t> d_esc8_pos; DevBuf<int32_t> d_esc8_val;
    int64_t n_esc8_col = -1, n_esc8_cnt = -1, esc8_col_cap = 0;    // escapes of the last pc_csr_get8
    int64_t n_rows = 0, nnz = 0;
    bool built = false, tfidf_done = false, scaled_done = false;
    DevBuf<float> d_scaled; DevBuf<double> d_colstat, d_rowscale;   // N3: StandardScaler(with_mean=False) values

    // tunables (pc_tune)
    int count_batch = 8, count_minb = 2, count_mode = 0;
    double table_load = 0.6;
    int part_mode = -1;                         // -1 auto, 0 direct global table, 1 partitioned
    int64_t part_bytes = 64ll << 20;            // target size of one L2-resident sub-table
    int probe_mode = -1;                        // -1 auto (stream when available), 0 roller, 1 stream
    int probe_batch = 4;
    int probe_roll = 1;                         // roller probe: 1 = filtered kernels of kernels_skm.cuh, 0 = the plain round-1 kernel
    int probe_threads = 0;                      // filtered probe: 0 auto, 512 (two CTAs per SM) or 1024
    // probe_filter: -1 auto, 0 off, 1 shared-memory bitmap, 2 L2-resident bitmap
    int probe_filter = -1;                      // shared-memory bitmap filter in front of the repeat table: -1 auto, 0 off, 1 on
    int64_t filter_bits = 1310720;              // 160 KB of shared memory (measured best: 1.0 M bits 3.58 ms, 1.31 M 3.49, 1.74 M 3.89 -- the
                                                // largest bitmap leaves the SM no L1 for the chunk / row look-ups)
    DevBuf<unsigned int> rfilter; unsigned int rfilter_bits = 0, gfilter_bits = 0;   // shared-memory / L2-resident bitmap
    int rep_load_inv = 0;                       // repeat table capacity = rep_load_inv * n_selected (0 = auto, see build_rep_table)
    bool stream_valid = false; int stream_k = 0;
    int part_blocks = 0;                        // blocks per bucket in k_count_part (0 = from the bucket size)

    // partitioned count state
    int pb = 0; uint64_t part_S = 0;
    DevBuf<uint64_t> kbuf, kstream; DevBuf<unsigned long long> d_part_hist; DevBuf<int64_t> d_seg;
    std::vector<int64_t> h_bucket_off;
    // measured on B200 (profiles/r01_kbench_count.md): the bucket count is bound by L2 sector operations; CAS-first
    // inserts (one ATOM per new k-mer, no load) are slower than load-first; HyperLogLog-sized denser tables and
    // in-kernel zeroing of the next sub-table were slower too and have been removed again
    int part_casfirst = 0, part_minb = 1;
    // threshold watch (pc_watch_threshold): k-mers that reach watch_low during a partitioned count
    uint32_t watch_low = 0; bool cross_valid = false; uint32_t cross_low = 0;
    DevBuf<uint64_t> cross_keys; DevBuf<unsigned long long> d_cross_n;
    int count_items = 2;                        // k-mers in flight per thread in the bucket kernel (2: 48 regs, best measured)

    // shared-memory count (part_mode 2, kernels_smem.cuh): level-2 buckets, then a compact (k-mer, count) list of the
    // k-mers with count >= list_min instead of a global table.  count_kind says what pc_select / pc_dump_counts read.
    int count_kind = 0;                         // 0 global table, 1 compact list
    uint32_t keep_min = 1;                      // tunable: lowest count worth keeping (the reference dumps mincount=3)
    int64_t smem_target = 0;                    // tunable: mean k-mer instances per sub-bucket (0 = half the table slots: an all-distinct sub-bucket loads its table to 0.5)
    int smem_slots() const { return (smem_variant == 1 || smem_variant == 5) ? 4096 : smem_variant == 2 ? 16384 : 8192; }
    int64_t smem_mean() const { return smem_target > 0 ? smem_target : (int64_t)smem_slots() / 2; }
    bool kbuf_shared = false;                   // the bucket buffer has been exported to peer processes (pc_kbuf_export)
    int rowsort_warps = 16;                     // tunable: warps per CTA in the per-row sort (8 or 16)
    int relief_path = 0;                        // set while pc_count runs: 1 = k-mer partition, 2 = super-k-mer records (see relieve_pressure)
    bool rec_shared = false;                    // the record buffer has been exported to peer processes (pc_skm_export)
    int select_raw = 0;                         // tunable: pc_select leaves the survivors unsorted and builds no repeat table (multi-GPU owners)
    bool sel_unsorted = false;
    int gram_preround = 0;                      // tunable: pc_densify_panel rounds the values to TF32, pc_gram_accumulate does not round again
    int gram_variant = 1;                       // tunable: Gram kernel, 0 = one CTA per 128 x 128 tile, 1 = persistent 128 x 256 tiles with two TMEM accumulators
    int csr_mode = -1;                          // tunable: rows -> CSR by -1 auto, 0 per-row radix sort, 1 shared-memory column bitmap (kernels_csr.cuh)
    int64_t csr_range_bits = 0, csr_ncap = 0;   // tunables (tests): columns per bitmap range / shared count slots (0 = as many as fit)
    DevBuf<unsigned long long> d_row_state;
    int scatter_pack = 1;                       // tunable: carry the bucket id in the spare top bits of the k-mer word (k <= 26)
    int scatter_tile = 0;                       // tunable: k-mers per tile of the level-1 scatter (0 auto, 8192: 2 CTAs/SM, 4096: 4 CTAs/SM)
    int smem_pb = 0;                            // tunable: level-1 bucket bits of the shared-memory count (0 = auto)
    int scatter_variant = 0;                    // tunable: 0 = two shared atomics per k-mer (256 x 32), 1 = one returning atomic (512 x 16)
    int smem_variant = 0;                       // tunable: 0 = 8192 slots x 512 threads x 4 k-mers in flight; 1 = 4096 x 256 x 4; 2 = 16384 x 1024 x 4; 3 = 8192 x 256 x 8; 4 = as 0 with read-first pair probing (measured slower: 6.1 vs 4.0 ms)
    DevBuf<uint64_t> kbuf2, list_keys; DevBuf<uint32_t> list_counts; DevBuf<unsigned long long> d_sub; DevBuf<int64_t> d_tile;
    DevBuf<unsigned int> d_ovf;
    int64_t n_list = 0, n_sub = 0, n_ovf = 0; uint32_t list_min = 1; unsigned int sub_nb2 = 0;
    // multi-GPU: level-2 histogram computed by the senders (pc_sub_hist), summed by the owner and staged for the next count
    const unsigned long long* staged_hist = nullptr; int64_t staged_n = 0; unsigned int staged_nb2 = 0;

    // super-k-mer count (part_mode 3, kernels_skm.cuh): the flat record stream of the extraction (later the level-2
    // output), the sub-bucket histogram ((k-mers << 32) | records per sub-bucket), its scan, the scatter cursors
    DevBuf<ulonglong2> rec_a, rec_b; DevBuf<unsigned long long> d_skm_hist, d_skm_off, d_skm_cur, d_skm_n;
    int64_t skm_cap_records = 0, skm_n_records = 0, skm_n_kmers = 0;
    int skm_m = 0, skm_w = 0, skm_nmax_used = 0;
    int skm_nmax = 16;                          // tunable: longest run (k-mers) one record carries (lane balance of the count kernel)
    int skm_expand = 0;                         // tunable: 1 = expand the records into 8-byte k-mers (k_rec_expand) and count them with k_count_smem; 0 = expand inside the count kernel (k_count_skm)
    DevBuf<unsigned long long> d_skm_koff;      // skm_expand 1: k-mer offsets of the sub-buckets + expansion cursors
    int skm_grp = 2;
    int skm_m_force = 0;                        // tunable: minimizer length (0 = skm_choose_m)                            // tunable: k-mers of a record in flight per thread in the count kernel (2, 4, 8)
    int skm_min_k = 17;                         // auto mode: shortest k that takes the super-k-mer path
    int skm_stage = 2560;                       // tunable: records staged per CTA of the extraction (x 16 bytes of shared memory)
    bool skm_used = false;
    uint64_t max_unit = 0;                      // k-mer instances of the largest counting unit (sub-bucket / bucket / table) of the last count

    // N1 subgenome-extraction session (pc_sub_*): per-subgenome dictionaries of the k being compared, then per k the
    // differential sets and their union mask table
    struct SubDict { DevBuf<pc::RepSlot> table; uint64_t cap = 0; DevBuf<uint64_t> keys; DevBuf<uint32_t> counts; int64_t n = 0; bool set = false; };
    struct SubLen { int k = 0; DevBuf<pc::RepSlot> mask; uint64_t cap = 0; DevBuf<uint64_t> diff_keys[pc::kMaxGroups];
                    DevBuf<uint32_t> diff_counts[pc::kMaxGroups]; int64_t n_diff[pc::kMaxGroups] = {}; };
    int sub_groups = 0, sub_k = 0, sub_n_len = 0;
    SubDict sub_dict[pc::kMaxGroups];
    SubLen sub_len[pc::kMaxSubK];
    DevBuf<uint64_t> sub_tmp_a, sub_tmp_b; DevBuf<int> d_cov;
    DevBuf<uint64_t> union_keys; DevBuf<uint32_t> union_counts; int64_t n_union_kept = -1;   // N4 (pc_sub_union)
    void sub_clear() {
        for (auto& d : sub_dict) { d.table.release(); d.keys.release(); d.counts.release(); d.cap = 0; d.n = 0; d.set = false; }
        for (auto& l : sub_len) {
            l.mask.release(); l.cap = 0; l.k = 0;
            for (int g = 0; g < pc::kMaxGroups; ++g) { l.diff_keys[g].release(); l.diff_counts[g].release(); l.n_diff[g] = 0; }
        }
        sub_tmp_a.release(); sub_tmp_b.release(); d_cov.release();
        union_keys.release(); union_counts.release(); n_union_kept = -1;
        sub_groups = sub_k = sub_n_len = 0;
    }
};

namespace {

thread_local pc_ctx* g_relief_ctx = nullptr;
bool relieve_pressure() {
    pc_ctx* c = g_relief_ctx;
    if (!c || c->relief_path == 0) return false;
    const int path = c->relief_path;
    c->relief_path = 0;                                        // once per count
    cudaStreamSynchronize(c->stream);
    // a new count invalidates the previous pass's repeat set and matrix (pc_count clears selected / built first)
    c->hits_a.release(); c->hits_b.release(); c->d_indices.release(); c->d_data.release(); c->d_tfidf.release(); c->d_row_state.release();
    c->d_data16.release(); c->d_dcol8.release(); c->d_cnt8.release(); c->d_esc_pos.release(); c->d_esc_val.release();
    c->d_esc8_pos.release(); c->d_esc8_val.release(); c->esc8_col_cap = 0; c->d_scaled.release();
    c->rtable.release(); c->rfp.release(); c->rfilter.release(); c->rtable8.release(); c->sel_a.release(); c->sel_b.release(); c->sel = nullptr;
    // the partition buffers of the path that is NOT running (never buffers peers may have mapped)
    if (path == 1 && !c->rec_shared) { c->rec_a.release(); c->rec_b.release(); c->skm_cap_records = 0; }
    if (path == 2) { c->kstream.release(); if (!c->kbuf_shared) { c->kbuf.release(); c->kbuf2.release(); } }
    return true;
}
struct ReliefScope {                                           // arms the relief for the duration of a pc_count
    pc_ctx* c;
    explicit ReliefScope(pc_ctx* c_) : c(c_) { g_relief_ctx = c; }
    ~ReliefScope() { c->relief_path = 0; g_relief_ctx = nullptr; }
};

struct StageTimer {
    pc_ctx* c; int st;
    StageTimer(pc_ctx* c_, int st_) : c(c_), st(st_) { cudaEventRecord(c->ev[st][0], c->stream); }
    ~StageTimer() { cudaEventRecord(c->ev[st][1], c->stream); c->ev_used[st] = true; }
};

#define LAUNCH(ctx, kernel, grid, block, ...)                                                      \
    do {                                                                                           \
        kernel<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__);                                \
        (ctx)->launches++;                                                                         \
        CU(cudaGetLastError());                                                                    \
    } while (0)

// Small device -> host read-backs (counters, bucket offsets) do not use cudaMemcpy: a copy engine serves its queue in
// order, so behind the bulk D2H of a neighbouring pipeline step (overlap.py) every 8-byte read waited for hundreds of
// megabytes and the steps serialised (measured: 35 ms instead of 20 ms per step with kernel times unchanged).  A tiny
// kernel writes the value into a page-locked, device-mapped mailbox instead; sync_stream() hands it to the caller.
__global__ void k_mail(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, size_t n_words) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

int read_small(pc_ctx* c, void* dst, const void* src_dev, size_t bytes) {
    if (bytes == 0) return PC_OK;
    if (!c->mailbox || (bytes & 3) || ((uintptr_t)src_dev & 3) || bytes > c->mailbox_cap - c->mailbox_used) {
        CU(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, c->stream));
        return PC_OK;
    }
    size_t off = c->mailbox_used;
    size_t words = bytes / 4;
    k_mail<<<(unsigned int)std::min<size_t>(64, (words + 255) / 256), 256, 0, c->stream>>>(
        static_cast<const uint32_t*>(src_dev), reinterpret_cast<uint32_t*>(c->mailbox + off), words);
    CU(cudaGetLastError());
    c->pending_reads.push_back({dst, off, bytes});
    c->mailbox_used = (off + bytes + 15) & ~(size_t)15;
    return PC_OK;
}

int sync_stream(pc_ctx* c) {
    CU(cudaStreamSynchronize(c->stream));
    for (const auto& r : c->pending_reads) memcpy(r.dst, c->mailbox + r.off, r.bytes);
    c->pending_reads.clear();
    c->mailbox_used = 0;
    return PC_OK;
}
#define SYNC(ctx) do { int rc_sync_ = sync_stream(ctx); if (rc_sync_) return rc_sync_; } while (0)
#define READ_SMALL(ctx, dst, src, bytes) do { int rc_rd_ = read_small(ctx, dst, src, bytes); if (rc_rd_) return rc_rd_; } while (0)

inline unsigned int blocks_for(int64_t n, int per) { return (unsigned int)std::max<int64_t>(1, (n + per - 1) / per); }

int use_device(pc_ctx* c) {
    if (!c) return fail(PC_EINVAL, "null context");
    CU(cudaSetDevice(c->device));
    c->pending_reads.clear(); c->mailbox_used = 0;            // a call that failed half-way leaves no stale read-backs
    return PC_OK;
}

int copy_out(pc_ctx* c, void* dst, const void* src_dev, size_t bytes, int loc) {
    if (bytes == 0) return PC_OK;
    CU(cudaMemcpyAsync(dst, src_dev, bytes, loc == PC_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    if (loc == PC_HOST) SYNC(c);
    return PC_OK;
}

// stable LSD radix sort of n uint64 keys over `bits` low bits; result pointer returned in *out (a or b)
constexpr int RS_WARPS = 8;
int radix_sort_u64(pc_ctx* c, uint64_t* a, uint64_t* b, int64_t n, int bits, uint64_t** out) {
    *out = a;
    if (n <= 1) return PC_OK;
    int n_ctas = (int)std::min<int64_t>(296, std::max<int64_t>(1, n / (RS_WARPS * 32 * 64)));   // >= 2048 keys per warp
    int NW = n_ctas * RS_WARPS;
    int64_t ipw = (((n + NW - 1) / NW) + 31) & ~(int64_t)31;
    int rc = c->rs_hist.alloc((size_t)256 * NW + 1);
    if (rc) return rc;
    uint64_t* src = a; uint64_t* dst = b;
    for (int shift = 0; shift < bits; shift += 8) {
        LAUNCH(c, (pc::k_rs_hist<uint64_t, RS_WARPS>), n_ctas, RS_WARPS * 32, src, n, ipw, shift, c->rs_hist.p, NW);
        LAUNCH(c, pc::k_rs_scan<unsigned int>, 1, 1024, c->rs_hist.p, (int64_t)256 * NW);
        LAUNCH(c, (pc::k_rs_scatter<uint64_t, RS_WARPS, false>), n_ctas, RS_WARPS * 32, src, dst,
               (const uint32_t*)nullptr, (uint32_t*)nullptr, n, ipw, shift, c->rs_hist.p, NW);
        std::swap(src, dst);
    }
    *out = src;
    return PC_OK;
}

int build_rep_table(pc_ctx* c) {
    StageTimer t(c, ST_REPTABLE);
    // load 1/8 while the table stays L2 resident (<= 96 MB), 1/4 beyond: the chain walk -- dependent L2 reads in a
    // divergent loop -- is what the probe pays for a fuller table (cfg2, no filter: load 1/2 8.1 ms, 1/4 4.1, 1/8 3.6)
    const int inv = c->rep_load_inv > 0 ? std::max(2, c->rep_load_inv)
                                        : ((uint64_t)c->n_sel * 8 * sizeof(pc::RepSlot) <= (96ull << 20) ? 8 : 4);
    c->rcap = std::max<uint64_t>(64, (uint64_t)c->n_sel * (uint64_t)inv);
    if (c->rcap >= (1ull << 32)) return fail(PC_EINVAL, "repeat set of %lld k-mers is too large for the repeat table", (long long)c->n_sel);
    int rc = c->rtable.alloc(c->rcap);
    if (rc) return rc;
    rc = c->rfp.alloc(c->rcap + 8);
    if (rc) return rc;
    CU(cudaMemsetAsync(c->rtable.p, 0, c->rcap * sizeof(pc::RepSlot), c->stream));
    CU(cudaMemsetAsync(c->rfp.p, 0, (c->rcap + 8) * sizeof(unsigned short), c->stream));
    if (c->n_sel > 0)
        LAUNCH(c, pc::k_rep_build, blocks_for(c->n_sel, 256), 256, c->sel, c->n_sel, c->rtable.p, c->rcap, c->rfp.p);
    // shared-memory filter of the probe: worth it while the bitmap stays sparse (>= 1.2 bits per repeat k-mer: fill < 57 %)
    c->rfilter_bits = 0; c->gfilter_bits = 0;
    // (the roller probe -- no k-mer stream -- measured better with the L2 bitmap than with the shared-memory one at every size:
    // 3.79 vs 3.91 ms at 673 k repeat k-mers; the shared-memory bitmap stays the choice of the stream probe)
    const bool smem_filter = c->n_sel > 0 && c->probe_filter != 0 && c->probe_filter != 2 &&
                             (c->probe_filter == 1 || (c->stream_valid && c->filter_bits * 5 >= c->n_sel * 6));
    // (not in front of the fingerprint walk, which already keeps misses away from the table: measured on 8 GPUs with
    // 5.6 M repeat k-mers, probe 7.4 ms with fingerprints alone, 8.8 ms with the bitmap in front of them)
    // (stream probe only; the roller probe measured best with the L2 bitmap in front of the plain 16-byte table at every size:
    // 6.8 M repeat k-mers, 438 MB table: 5.1 ms, against 7.9 with the fingerprint walk and 12.1 with the compact 8-byte table)
    const bool fp_walk = c->probe_fp == 1 || (c->probe_fp < 0 && c->stream_valid && c->rcap * sizeof(pc::RepSlot) > (size_t)(256ull << 20));
    if (c->n_sel > 0 && !smem_filter && (c->probe_filter == 2 || (c->probe_filter < 0 && !fp_walk))) {
        // L2-resident bitmap, 16 bits per repeat k-mer (false positives ~6 %): for repeat sets too large for the
        // shared-memory filter (measured at 673 k repeat k-mers: no filter 5.16 ms, this 4.04 ms, shared-memory filter
        // 3.48 ms; at 2.2 M repeat k-mers / 4.5 Gbp, where the table is beyond L2: 52.9 -> 41.7 ms)
        uint64_t nb = std::min<uint64_t>((uint64_t)c->n_sel * 16, 1ull << 31);
        unsigned int nbits = (unsigned int)std::max<uint64_t>(1024, nb) & ~127u;
        rc = c->rfilter.alloc(nbits / 32 + 4); if (rc) return rc;
        CU(cudaMemsetAsync(c->rfilter.p, 0, ((size_t)nbits / 32 + 4) * sizeof(unsigned int), c->stream));
        LAUNCH(c, pc::k_filter_build, blocks_for(c->n_sel, 256), 256, c->sel, c->n_sel, nbits, c->rfilter.p);
        c->gfilter_bits = nbits;
    }
    c->rcap8 = 0;
    // compact 8-byte-slot table: opt-in only (pc_tune probe_table8 1).  Measured slower than the 16-byte slots with inline
    // keys both where everything fits L2 (673 k repeat k-mers: 5.78 ms against 3.79) and where the 16-byte table is 438 MB
    // (6.8 M: 12.1 against 5.1): the verification read of the key array and the longer probe chains at load 0.5 cost more
    // than the DRAM sectors of the hits, once the L2 bitmap keeps the misses away from the table
    if (c->n_sel > 0 && c->probe_table8 == 1) {
        c->rcap8 = std::max<uint64_t>(64, (uint64_t)c->n_sel * 2);
        rc = c->rtable8.alloc(c->rcap8); if (rc) return rc;
        CU(cudaMemsetAsync(c->rtable8.p, 0, c->rcap8 * sizeof(unsigned long long), c->stream));
        LAUNCH(c, pc::k_rep8_build, blocks_for(c->n_sel, 256), 256, c->sel, c->n_sel, c->rtable8.p, c->rcap8);
    }
    if (smem_filter) {
        unsigned int nbits = (unsigned int)std::min<int64_t>(std::max<int64_t>(c->filter_bits, 1024), 1744896) & ~127u;
        rc = c->rfilter.alloc(nbits / 32 + 4); if (rc) return rc;
        CU(cudaMemsetAsync(c->rfilter.p, 0, (nbits / 32 + 4) * sizeof(unsigned int), c->stream));
        LAUNCH(c, pc::k_filter_build, blocks_for(c->n_sel, 256), 256, c->sel, c->n_sel, nbits, c->rfilter.p);
        c->rfilter_bits = nbits;
    }
    return PC_OK;
}

}  // namespace

extern "C" {

int pc_version(void) { return 100; }
const char* pc_last_error(void) { return g_err.c_str(); }

int pc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int pc_create(int device, pc_ctx** out) {
    if (!out) return fail(PC_EINVAL, "out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(PC_ECUDA, "no CUDA device: polycracker_b200 has no CPU fallback (%s)", cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return fail(PC_EINVAL, "device %d out of range (have %d)", device, n);
    CU(cudaSetDevice(device));
    pc_ctx* c = new pc_ctx();
    c->device = device;
    CU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    c->stream = c->own_stream;
    for (int i = 0; i < ST_N; ++i) {
        CU(cudaEventCreate(&c->ev[i][0])); CU(cudaEventCreate(&c->ev[i][1])); c->ev_used[i] = false;
    }
    c->mailbox_cap = 256 << 10;
    if (cudaHostAlloc(reinterpret_cast<void**>(&c->mailbox), c->mailbox_cap, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError(); c->mailbox = nullptr; c->mailbox_cap = 0;          // read_small falls back to cudaMemcpyAsync
    }
    int rc = c->d_stats.alloc(1); if (rc) { delete c; return rc; }
    rc = c->d_counter.alloc(64); if (rc) { delete c; return rc; }
    *out = c;
    return PC_OK;
}

int pc_destroy(pc_ctx* c) {
    if (!c) return PC_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->ascii.release(); c->d_word0.release(); c->d_begin.release(); c->d_len.release();
    c->words.release(); c->nmask.release(); c->table_keys.release(); c->table_counts.release(); c->d_stats.release();
    c->kbuf.release(); c->kstream.release(); c->d_part_hist.release(); c->d_seg.release(); c->cross_keys.release(); c->d_cross_n.release();
    c->d_scaled.release(); c->d_colstat.release(); c->d_rowscale.release();
    c->kbuf2.release(); c->list_keys.release(); c->list_counts.release(); c->d_sub.release(); c->d_tile.release(); c->d_ovf.release();
    c->rec_a.release(); c->rec_b.release(); c->d_skm_hist.release(); c->d_skm_off.release(); c->d_skm_cur.release(); c->d_skm_n.release(); c->d_skm_koff.release();
    c->sel_a.release(); c->sel_b.release(); c->rs_hist.release(); c->d_counter.release(); c->rtable.release(); c->rfp.release(); c->rfilter.release(); c->rtable8.release();
    c->d_chunk_row.release(); c->d_row_base.release(); c->d_row_cursor.release(); c->hits_a.release();
    c->hits_b.release(); c->d_row_nnz.release(); c->d_row_state.release(); c->d_indptr.release(); c->d_indices.release();
    c->d_data.release(); c->d_idf.release(); c->d_tfidf.release(); c->d_df.release(); c->d_df_global.release();
    c->d_row_norm.release(); c->d_data16.release(); c->d_esc_pos.release(); c->d_esc_val.release();
    c->d_dcol8.release(); c->d_cnt8.release(); c->d_esc8_pos.release(); c->d_esc8_val.release();
    c->sub_clear();
    for (int i = 0; i < ST_N; ++i) { cudaEventDestroy(c->ev[i][0]); cudaEventDestroy(c->ev[i][1]); }
    if (c->mailbox) cudaFreeHost(c->mailbox);
    cudaStreamDestroy(c->own_stream);
    cudaStreamDestroy(c->aux_stream);
    cudaEventDestroy(c->ev_fork); cudaEventDestroy(c->ev_join);
    delete c;
    return PC_OK;
}

int pc_set_stream(pc_ctx* c, void* s) {
    if (!c) return fail(PC_EINVAL, "null context");
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return PC_OK;
}

int pc_sync(pc_ctx* c) {
    int rc = use_device(c); if (rc) return rc;
    SYNC(c);
    return PC_OK;
}

// ---------------------------------------------------------------------------------------------- a1
int64_t pc_chunk_count(const int64_t* seq_len, int64_t n_seqs, int64_t C) {
    if (!seq_len || C <= 0) return -1;
    int64_t n = 0;
    for (int64_t s = 0; s < n_seqs; ++s) {
        int64_t L = seq_len[s];
        if (L <= 0) continue;
        n += (L + C - 1) / C;                 // len(arange(0, L, C)) chunks: pairs + tail
    }
    return n;
}

int pc_chunk_table(const int64_t* seq_len, int64_t n_seqs, int64_t C, int64_t* chunk_seq, int64_t* chunk_start,
                   int64_t* chunk_end) {
    if (!seq_len || !chunk_seq || !chunk_start || !chunk_end || C <= 0) return fail(PC_EINVAL, "bad chunk_table arguments");
    int64_t n = 0;
    for (int64_t s = 0; s < n_seqs; ++s) {
        int64_t L = seq_len[s];
        if (L <= 0) continue;
        for (int64_t p = 0; p < L; p += C) {
            chunk_seq[n] = s; chunk_start[n] = p; chunk_end[n] = std::min(L, p + C); ++n;
        }
    }
    return PC_OK;
}

// ---------------------------------------------------------------------------------------------- K1
int pc_ascii_buffer(pc_ctx* c, int64_t n_bytes, uint8_t** dev_ptr) {
    int rc = use_device(c); if (rc) return rc;
    if (n_bytes < 0 || !dev_ptr) return fail(PC_EINVAL, "bad ascii buffer request");
    int64_t padded = ((n_bytes + 255) & ~(int64_t)255) + 256;
    rc = c->ascii.alloc((size_t)padded); if (rc) return rc;
    c->ascii_padded = padded;
    *dev_ptr = c->ascii.p;
    return PC_OK;
}

int pc_load_sequences(pc_ctx* c, const uint8_t* ascii, int64_t n_bytes, int loc, const int64_t* chunk_begin,
                      const int64_t* chunk_end, int64_t n_chunks) {
    int rc = use_device(c); if (rc) return rc;
    if (n_bytes < 0 || n_chunks < 0 || (n_bytes > 0 && !ascii) || (n_chunks > 0 && (!chunk_begin || !chunk_end)))
        return fail(PC_EINVAL, "bad load_sequences arguments");
    c->loaded = c->counted = c->selected = c->built = c->tfidf_done = false;
    c->stream_valid = false;
    c->h_word0.assign(n_chunks + 1, 0); c->h_begin.assign(std::max<int64_t>(n_chunks, 1), 0); c->h_len.assign(std::max<int64_t>(n_chunks, 1), 0);
    int64_t prev_end = 0, wsum = 0;
    for (int64_t i = 0; i < n_chunks; ++i) {
        int64_t b = chunk_begin[i], e = chunk_end[i];
        if (b < prev_end || e < b || e > n_bytes) return fail(PC_EINVAL, "chunk %lld [%lld,%lld) is out of order or out of range", (long long)i, (long long)b, (long long)e);
        prev_end = e;
        c->h_begin[i] = b; c->h_len[i] = e - b; c->h_word0[i] = wsum;
        wsum += (e - b) / 32 + 1;             // >= 1 padding base after every chunk
    }
    c->h_word0[n_chunks] = wsum;
    c->n_chunks = n_chunks; c->n_words = wsum;

    const uint8_t* src_dev = nullptr; int64_t src_padded = 0;
    {
        StageTimer t(c, ST_H2D);
        if (loc == PC_HOST) {
            uint8_t* dptr = nullptr;
            rc = pc_ascii_buffer(c, n_bytes, &dptr); if (rc) return rc;
            if (n_bytes) CU(cudaMemcpyAsync(dptr, ascii, (size_t)n_bytes, cudaMemcpyHostToDevice, c->stream));
            src_dev = dptr; src_padded = c->ascii_padded;
        } else if (ascii == c->ascii.p && c->ascii.p) {
            src_dev = c->ascii.p; src_padded = c->ascii_padded;
        } else {
            src_dev = ascii; src_padded = n_bytes;                // caller-owned: no slack assumed
        }
        rc = c->d_word0.alloc(n_chunks + 1); if (rc) return rc;
        rc = c->d_begin.alloc(n_chunks); if (rc) return rc;
        rc = c->d_len.alloc(n_chunks); if (rc) return rc;
        CU(cudaMemcpyAsync(c->d_word0.p, c->h_word0.data(), (n_chunks + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        if (n_chunks) {
            CU(cudaMemcpyAsync(c->d_begin.p, c->h_begin.data(), n_chunks * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemcpyAsync(c->d_len.p, c->h_len.data(), n_chunks * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        }
    }
    rc = c->words.alloc(wsum + 1); if (rc) return rc;
    rc = c->nmask.alloc(wsum + 1); if (rc) return rc;
    {
        StageTimer t(c, ST_PACK);
        LAUNCH(c, pc::k_pack, blocks_for(wsum + 1, 256), 256, src_dev, src_padded, c->d_word0.p, c->d_begin.p,
               c->d_len.p, n_chunks, wsum + 1, c->words.p, c->nmask.p);
    }
    // the h_* vectors are re-used by later host code; the async copies above read them now
    SYNC(c);
    c->loaded = true;
    return PC_OK;
}

int pc_packed_info(pc_ctx* c, int64_t* n_words, int64_t* n_chunks) {
    if (!c || !c->loaded) return fail(PC_ESTATE, "no sequences loaded");
    if (n_words) *n_words = c->n_words;
    if (n_chunks) *n_chunks = c->n_chunks;
    return PC_OK;
}

int pc_packed_get(pc_ctx* c, uint64_t* words, uint32_t* nmask, int64_t* chunk_word0, int loc) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->loaded) return fail(PC_ESTATE, "no sequences loaded");
    if (words) { rc = copy_out(c, words, c->words.p, c->n_words * sizeof(uint64_t), loc); if (rc) return rc; }
    if (nmask) { rc = copy_out(c, nmask, c->nmask.p, c->n_words * sizeof(uint32_t), loc); if (rc) return rc; }
    if (chunk_word0) { rc = copy_out(c, chunk_word0, c->d_word0.p, (c->n_chunks + 1) * sizeof(int64_t), loc); if (rc) return rc; }
    return PC_OK;
}

// ---------------------------------------------------------------------------------------------- K3
static int table_alloc(pc_ctx* c, int k, int64_t capacity) {
    if (k < 1 || k > 32) return fail(PC_EINVAL, "k = %d unsupported: this path handles 1 <= k <= 32 (uint64 k-mers)", k);
    // (no driver query when the table does not have to grow: this runs in the middle of every step)
    if (c->table_keys.p && c->table_counts.p && c->table_keys.n >= (size_t)capacity && c->table_counts.n >= (size_t)capacity) return PC_OK;
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    size_t need = (size_t)capacity * (size_t)PC_SLOT_BYTES;
    if (need > free_b + c->table_keys.n * (size_t)PC_SLOT_BYTES)
        return fail(PC_ENOMEM, "count table of %lld slots (%.1f GB) does not fit in free HBM (%.1f GB): shard the genome over more GPUs",
                    (long long)capacity, need / 1e9, free_b / 1e9);
    int rc = c->table_keys.alloc((size_t)capacity); if (rc) return rc;
    return c->table_counts.alloc((size_t)capacity);
}

// Start zeroing `capacity` slots on the side stream (ordered after everything already queued on the main stream), so
// the 16 B/slot memset overlaps the extract + scatter kernels, which leave most of the HBM bandwidth unused.
static int table_prezero(pc_ctx* c, int k, int64_t capacity) {
    if (!c->prezero) return PC_OK;
    if (capacity < 64) capacity = 64;
    int rc = table_alloc(c, k, capacity); if (rc) return rc;
    CU(cudaEventRecord(c->ev_fork, c->stream));
    CU(cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
    CU(cudaEventRecord(c->ev[ST_TABLE_INIT][0], c->aux_stream));
    CU(cudaMemsetAsync(c->table_keys.p, 0, (size_t)capacity * 8, c->aux_stream));
    CU(cudaMemsetAsync(c->table_counts.p, 0, (size_t)capacity * 4, c->aux_stream));
    CU(cudaEventRecord(c->ev[ST_TABLE_INIT][1], c->aux_stream));
    c->ev_used[ST_TABLE_INIT] = true;
    CU(cudaEventRecord(c->ev_join, c->aux_stream));
    c->zeroed_slots = (uint64_t)capacity;
    return PC_OK;
}

static int table_prepare(pc_ctx* c, int k, int64_t capacity, bool reset_stats = true) {
    if (capacity < 64) capacity = 64;
    if (c->zeroed_slots > 0 && c->table_keys.p && c->table_keys.n >= (size_t)capacity && c->table_counts.n >= (size_t)capacity) {
        // a side-stream memset is pending: join it, zero whatever it did not cover
        CU(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        if ((uint64_t)capacity > c->zeroed_slots)
        {
            CU(cudaMemsetAsync(c->table_keys.p + c->zeroed_slots, 0, ((size_t)capacity - c->zeroed_slots) * 8, c->stream));
            CU(cudaMemsetAsync(c->table_counts.p + c->zeroed_slots, 0, ((size_t)capacity - c->zeroed_slots) * 4, c->stream));
        }
        c->zeroed_slots = 0;
    } else {
        if (c->zeroed_slots > 0) { CU(cudaStreamWaitEvent(c->stream, c->ev_join, 0)); c->zeroed_slots = 0; }
        int rc = table_alloc(c, k, capacity); if (rc) return rc;
        StageTimer t(c, ST_TABLE_INIT);
        CU(cudaMemsetAsync(c->table_keys.p, 0, (size_t)capacity * 8, c->stream));
        CU(cudaMemsetAsync(c->table_counts.p, 0, (size_t)capacity * 4, c->stream));
    }
    c->capacity = (uint64_t)capacity; c->k = k;
    c->cross_valid = false;
    if (reset_stats) CU(cudaMemsetAsync(c->d_stats.p, 0, sizeof(pc::CountStats), c->stream));
    return PC_OK;
}

// counts are uint32: a counting unit (one shared-memory table, one sub-table, the whole direct table) that received 2^32 or
// more k-mer instances could have wrapped a counter.  That needs one k-mer with > 4.29e9 occurrences on one GPU (a
// multi-Gbp homopolymer); it is refused loudly instead of returning a wrapped count ("counts saturate / flag").
static int check_unit_size(pc_ctx* c, uint64_t instances) {
    c->max_unit = std::max<uint64_t>(c->max_unit, instances);
    if (instances >= (1ull << 32))
        return fail(PC_ENOMEM, "a counting unit holds %llu k-mer instances: a uint32 counter could wrap -- shard the genome over "
                    "more GPUs (counts are never returned wrapped)", (unsigned long long)instances);
    return PC_OK;
}

static int fetch_stats(pc_ctx* c) {
    READ_SMALL(c, &c->h_stats, c->d_stats.p, sizeof(pc::CountStats));
    SYNC(c);
    if (c->h_stats.overflow)
        return fail(PC_ENOMEM, "count table overflow: %llu insertions hit the probe limit (capacity %llu)",
                    c->h_stats.overflow, (unsigned long long)c->capacity);
    return PC_OK;
}

#define COUNT_CASE(B, M, MODE)                                                                               \
    if (c->count_batch == B && c->count_minb == M && c->count_mode == MODE) {                                \
        LAUNCH(c, (pc::k_count<B, M, MODE>), blocks_for(c->n_words, 256), 256, c->words.p, c->nmask.p,       \
               c->n_words, c->k, c->table(), c->capacity, c->d_stats.p);                                      \
        return PC_OK;                                                                                        \
    }
static int launch_count(pc_ctx* c) {
    COUNT_CASE(8, 2, 0) COUNT_CASE(8, 3, 0) COUNT_CASE(8, 4, 0) COUNT_CASE(16, 2, 0) COUNT_CASE(4, 4, 0)
    COUNT_CASE(8, 2, 1) COUNT_CASE(8, 3, 1) COUNT_CASE(8, 4, 1) COUNT_CASE(16, 2, 1) COUNT_CASE(4, 4, 1)
    COUNT_CASE(8, 2, 2) COUNT_CASE(8, 3, 2) COUNT_CASE(8, 4, 2) COUNT_CASE(16, 2, 2) COUNT_CASE(4, 4, 2)
    return fail(PC_EINVAL, "no k_count instantiation for batch=%d minb=%d mode=%d", c->count_batch, c->count_minb, c->count_mode);
}

// extract (k-mer stream + bucket histogram) and tile scatter; leaves the bucket-sorted k-mers in c->kbuf and the
// bucket offsets in c->h_bucket_off.  d_stats must have been reset by the caller.
static int do_partition(pc_ctx* c, int k, int pb) {
    int rc;
    const uint64_t P = 1ull << pb;
    int64_t windows = 0;
    for (int64_t i = 0; i < c->n_chunks; ++i) windows += std::max<int64_t>(0, c->h_len[i] - k + 1);
    int64_t n_stream = ((c->n_words + 255) / 256) * 256 * 32;              // one slot per window start
    rc = c->kstream.alloc((size_t)std::max<int64_t>(n_stream, 1)); if (rc) return rc;
    rc = c->kbuf.alloc((size_t)std::max<int64_t>(windows, 1)); if (rc) return rc;
    rc = c->d_part_hist.alloc((size_t)3 * P + 2); if (rc) return rc;
    unsigned long long* bucket_hist = c->d_part_hist.p;
    unsigned long long* bucket_off = bucket_hist + P;                      // [P + 1]
    unsigned long long* bucket_cur = bucket_off + P + 1;                   // [P]
    {
        StageTimer t(c, ST_PART_HIST);
        CU(cudaMemsetAsync(bucket_hist, 0, P * sizeof(unsigned long long), c->stream));
        if (c->n_words > 0) {
            c->launches++;
            pc::k_extract<<<blocks_for(c->n_words, 256), 256, P * sizeof(unsigned int), c->stream>>>(
                c->words.p, c->nmask.p, c->n_words, k, pb, c->kstream.p, bucket_hist, c->d_stats.p);
            CU(cudaGetLastError());
        }
        LAUNCH(c, pc::k_bucket_scan, 1, 1024, bucket_hist, (unsigned int)P, bucket_off, bucket_cur);
    }
    c->stream_valid = c->n_words > 0; c->stream_k = k;
    if (c->n_words > 0) {
        StageTimer t(c, ST_PART_SCATTER);
        c->launches++;
        const bool pack = 2 * k <= PC_PACK_SHIFT && c->scatter_pack != 0;   // bucket id in the spare top bits of the k-mer word
#define TILE_SCATTER(TILE_, PACK_, PER_SM)                                                                                  \
        do {                                                                                                               \
            size_t smem = (size_t)(TILE_) * 8 + P * 8 + P * 4 + (P + 2) * 4 + (size_t)(TILE_) * 2 + 16;                    \
            CU(cudaFuncSetAttribute((pc::k_tile_scatter<TILE_, PACK_>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            unsigned int grid = (unsigned int)std::min<int64_t>(148 * (PER_SM), (n_stream + (TILE_) - 1) / (TILE_));       \
            pc::k_tile_scatter<TILE_, PACK_><<<grid, 256, smem, c->stream>>>(c->kstream.p, n_stream, pb, bucket_off, bucket_cur, c->kbuf.p); \
        } while (0)
        // 4096-k-mer tiles (4 CTAs per SM) while the runs stay >= 8 k-mers long; measured at P = 512: 2.38 ms (8192, two
        // arrays) -> 2.22 (bucket id packed) -> 2.10 (4096 + packed)
        // (not when the bucket buffer is exported to peers over CUDA IPC: measured on 2 GPUs 3.63 ms with the small tiles
        // against 2.58 ms with the large ones)
        const bool small_tile = c->scatter_tile == 4096 || (c->scatter_tile == 0 && P <= 512 && !c->kbuf_shared);
        if (small_tile) { if (pack) TILE_SCATTER(4096, true, 4); else TILE_SCATTER(4096, false, 4); }
        else { if (pack) TILE_SCATTER(8192, true, 2); else TILE_SCATTER(8192, false, 2); }
#undef TILE_SCATTER
        CU(cudaGetLastError());
    }
    c->h_bucket_off.assign(P + 1, 0);
    static_assert(sizeof(unsigned long long) == sizeof(int64_t), "offset width");
    READ_SMALL(c, c->h_bucket_off.data(), bucket_off, (P + 1) * sizeof(int64_t));
    READ_SMALL(c, &c->h_stats, c->d_stats.p, sizeof(pc::CountStats));
    SYNC(c);
    return PC_OK;
}

// count n_buckets buckets (sub-table i <- segments [i*n_seg, (i+1)*n_seg) of keys_dev) into a fresh table
static int count_buckets_table(pc_ctx* c, int k, int pb, int n_buckets, const uint64_t* keys_dev, const int64_t* seg_begin,
                               const int64_t* seg_len, int n_seg, int64_t capacity, const uint64_t* const* seg_base_host = nullptr) {
    int rc;
    int64_t total = 0, max_bucket = 0;
    for (int q = 0; q < n_buckets; ++q) {
        int64_t t = 0;
        for (int s = 0; s < n_seg; ++s) {
            if (seg_len[(int64_t)q * n_seg + s] < 0 || seg_begin[(int64_t)q * n_seg + s] < 0) return fail(PC_EINVAL, "bad segment");
            t += seg_len[(int64_t)q * n_seg + s];
        }
        total += t; max_bucket = std::max(max_bucket, t);
    }
    { int rc_u = check_unit_size(c, (uint64_t)max_bucket); if (rc_u) return rc_u; }
    if (capacity <= 0) {
        double space = k >= 31 ? 9.3e18 : (double)(1ull << (2 * k));
        capacity = (int64_t)(std::min((double)total, space) / c->table_load) + 1024;
    }
    uint64_t S = ((uint64_t)capacity + n_buckets - 1) / n_buckets;
    // a bucket fed by heavy-hitter k-mers may hold more instances than the average, never more distinct keys than
    // instances: keep every sub-table at least big enough for the mean and let the probe limit report the rest
    if (S < 64) S = 64;
    capacity = (int64_t)(S * (uint64_t)n_buckets);
    rc = table_prepare(c, k, capacity, /*reset_stats=*/false); if (rc) return rc;
    c->pb = pb; c->part_S = S;
    rc = c->d_seg.alloc((size_t)3 * n_buckets * n_seg); if (rc) return rc;
    CU(cudaMemcpyAsync(c->d_seg.p, seg_begin, (size_t)n_buckets * n_seg * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_seg.p + (size_t)n_buckets * n_seg, seg_len, (size_t)n_buckets * n_seg * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    const long long* d_beg = reinterpret_cast<const long long*>(c->d_seg.p);
    const long long* d_len = d_beg + (size_t)n_buckets * n_seg;
    const uint64_t* const* d_base = nullptr;
    if (seg_base_host) {
        static_assert(sizeof(uint64_t*) == sizeof(int64_t), "pointer width");
        CU(cudaMemcpyAsync(c->d_seg.p + (size_t)2 * n_buckets * n_seg, seg_base_host, (size_t)n_buckets * n_seg * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        d_base = reinterpret_cast<const uint64_t* const*>(c->d_seg.p + (size_t)2 * n_buckets * n_seg);
    }
    pc::Watch w{0u, nullptr, nullptr, 0ull};
    c->cross_valid = false;
    if (c->watch_low > 0) {
        unsigned long long cap = (unsigned long long)(total / c->watch_low) + 1024ull;
        rc = c->cross_keys.alloc((size_t)cap); if (rc) return rc;
        rc = c->d_cross_n.alloc(1); if (rc) return rc;
        CU(cudaMemsetAsync(c->d_cross_n.p, 0, sizeof(unsigned long long), c->stream));
        w = pc::Watch{c->watch_low, c->d_cross_n.p, c->cross_keys.p, cap};
        c->cross_valid = true; c->cross_low = c->watch_low;
    }
    if (total > 0) {
        StageTimer t(c, ST_COUNT);
        // blocks per bucket: ~14 k-mers per thread keeps the resident blocks on one or two buckets (L2) and still
        // amortises the block prologue (measured: B=512 best at 1.5 M keys/bucket, B=1024 at 3.9 M)
        unsigned int B = (unsigned int)c->part_blocks;
        if (c->part_blocks <= 0) {
            double want = (double)(total / n_buckets) / 3600.0;
            B = 64; while (B < 4096 && (double)B * 1.41 < want) B <<= 1;
        }
        unsigned int grid = (unsigned int)n_buckets * B;
#define COUNT_PART(I, CF, W, MB) LAUNCH(c, (pc::k_count_part<I, CF, W, MB>), grid, 256, keys_dev, d_base, d_beg, d_len, n_seg, pb, S, B, c->table(), c->d_stats.p, w)
        if (w.low) {
            if (c->part_casfirst) COUNT_PART(4, true, true, 1); else COUNT_PART(4, false, true, 1);
        } else if (c->part_casfirst) COUNT_PART(4, true, false, 1);
        else if (c->count_items == 2) {
            if (c->part_minb == 6) COUNT_PART(2, false, false, 6); else if (c->part_minb == 8) COUNT_PART(2, false, false, 8); else COUNT_PART(2, false, false, 1);
        } else if (c->count_items == 1) {
            if (c->part_minb == 8) COUNT_PART(1, false, false, 8); else COUNT_PART(1, false, false, 1);
        } else {
            if (c->part_minb == 5) COUNT_PART(4, false, false, 5); else if (c->part_minb == 6) COUNT_PART(4, false, false, 6); else COUNT_PART(4, false, false, 1);
        }
#undef COUNT_PART
    }
    rc = fetch_stats(c); if (rc) return rc;                   // synchronises: the host segment arrays may go away
    c->counted = true; c->count_kind = 0;
    return PC_OK;
}

// Shared-memory count (kernels_smem.cuh): level-2 histogram + scatter of the n_buckets level-1 buckets into sub-buckets
// of about smem_target k-mers, one CTA-private shared-memory table per sub-bucket, survivors (count >= keep_min) to a
// compact list.  Sub-buckets that do not fit their table are re-counted through a global table (count_buckets_table).
extern "C++" {
template <int SLOTS, int THREADS, int PER, bool PAIR>
static int launch_count_smem(pc_ctx* c, unsigned int n_sub, unsigned long long list_cap, unsigned long long* d_nlist, unsigned int* d_novf) {
    size_t smem = (size_t)SLOTS * 12;
    CU(cudaFuncSetAttribute(pc::k_count_smem<SLOTS, THREADS, PER, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = std::max(1, (int)((size_t)(220 << 10) / (smem + 1024)));
    per_sm = std::min(per_sm, 2048 / THREADS);
    unsigned int grid = std::min<unsigned int>(n_sub, 148u * (unsigned int)per_sm);
    c->launches++;
    pc::k_count_smem<SLOTS, THREADS, PER, PAIR><<<grid, THREADS, smem, c->stream>>>(
        c->kbuf2.p, c->d_sub.p, n_sub, c->keep_min, c->list_keys.p, c->list_counts.p, list_cap, d_nlist, c->d_ovf.p, d_novf, c->d_stats.p);
    CU(cudaGetLastError());
    return PC_OK;
}
}  // extern "C++"

static unsigned int sub_plan(const pc_ctx* c, int64_t total, int n_buckets) {
    int64_t nb2 = (total + (int64_t)n_buckets * c->smem_mean() - 1) / ((int64_t)n_buckets * c->smem_mean());
    return (unsigned int)std::min<int64_t>(PC_SUB_MAX, std::max<int64_t>(1, nb2));
}

static int count_buckets_smem(pc_ctx* c, int k, int pb, int n_buckets, const uint64_t* keys_dev, const int64_t* seg_begin,
                              const int64_t* seg_len, int n_seg, const uint64_t* const* seg_base_host) {
    int rc;
    constexpr int TILE = 8192;
    const int64_t n_pairs = (int64_t)n_buckets * n_seg;
    std::vector<int64_t> tile_start(n_pairs + 1, 0);
    int64_t total = 0;
    for (int64_t p = 0; p < n_pairs; ++p) {
        if (seg_len[p] < 0 || seg_begin[p] < 0) return fail(PC_EINVAL, "bad segment");
        total += seg_len[p];
        tile_start[p + 1] = tile_start[p] + (seg_len[p] + TILE - 1) / TILE;
    }
    if (c->keep_min < 1) c->keep_min = 1;
    const int64_t n_tiles = tile_start[n_pairs];
    unsigned int NB2 = sub_plan(c, total, n_buckets);
    const unsigned long long* staged = nullptr;
    if (c->staged_hist && c->staged_nb2 >= 1 && c->staged_nb2 <= PC_SUB_MAX && c->staged_n == (int64_t)n_buckets * c->staged_nb2) {
        staged = c->staged_hist; NB2 = c->staged_nb2;
    }
    c->staged_hist = nullptr; c->staged_n = 0; c->staged_nb2 = 0;  // one-shot
    const int64_t n_sub = (int64_t)n_buckets * NB2;
    if (n_sub >= (1ll << 31)) return fail(PC_EINVAL, "too many sub-buckets");
    c->k = k; c->pb = pb; c->sub_nb2 = NB2; c->n_sub = n_sub; c->cross_valid = false;
    c->part_S = 0; c->capacity = (uint64_t)n_sub * (uint64_t)c->smem_slots();

    rc = c->d_seg.alloc((size_t)3 * n_pairs); if (rc) return rc;
    rc = c->d_tile.alloc((size_t)n_pairs + 1); if (rc) return rc;
    rc = c->d_sub.alloc((size_t)2 * n_sub + 8); if (rc) return rc;
    rc = c->d_ovf.alloc((size_t)n_sub + 2); if (rc) return rc;
    rc = c->kbuf2.alloc((size_t)std::max<int64_t>(total, 1)); if (rc) return rc;
    const unsigned long long list_cap = (unsigned long long)(total / c->keep_min) + 1024ull;
    rc = c->list_keys.alloc((size_t)list_cap); if (rc) return rc;
    rc = c->list_counts.alloc((size_t)list_cap); if (rc) return rc;
    CU(cudaMemcpyAsync(c->d_seg.p, seg_begin, (size_t)n_pairs * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_seg.p + n_pairs, seg_len, (size_t)n_pairs * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    const int64_t* d_beg = c->d_seg.p;
    const int64_t* d_len = c->d_seg.p + n_pairs;
    const uint64_t* const* d_base = nullptr;
    if (seg_base_host) {
        CU(cudaMemcpyAsync(c->d_seg.p + 2 * n_pairs, seg_base_host, (size_t)n_pairs * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        d_base = reinterpret_cast<const uint64_t* const*>(c->d_seg.p + 2 * n_pairs);
    }
    CU(cudaMemcpyAsync(c->d_tile.p, tile_start.data(), (size_t)(n_pairs + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    unsigned long long* sub_off = c->d_sub.p;                              // [n_sub + 1] sizes, then offsets
    unsigned long long* sub_cur = c->d_sub.p + n_sub + 1;                  // [n_sub] scatter cursors
    unsigned long long* d_nlist = c->d_sub.p + 2 * n_sub + 2;              // [1]
    unsigned int* d_novf = reinterpret_cast<unsigned int*>(c->d_sub.p + 2 * n_sub + 3);
    CU(cudaMemsetAsync(c->d_sub.p, 0, ((size_t)2 * n_sub + 8) * sizeof(unsigned long long), c->stream));
    if (total > 0) {
        {
            StageTimer t(c, ST_PART2);
            unsigned int grid = (unsigned int)std::min<int64_t>(148 * 8, n_tiles);
            if (staged) {
                CU(cudaMemcpyAsync(sub_off, staged, (size_t)n_sub * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
            } else {
                c->launches++;
                pc::k_sub_hist<TILE><<<grid, 256, NB2 * sizeof(unsigned int), c->stream>>>(
                    keys_dev, d_base, d_beg, d_len, c->d_tile.p, (int)n_pairs, n_seg, pb, NB2, sub_off);
                CU(cudaGetLastError());
            }
            LAUNCH(c, pc::k_rs_scan<unsigned long long>, 1, 1024, sub_off, n_sub);
            LAUNCH(c, pc::k_max_diff, (unsigned int)std::min<int64_t>(148, (n_sub + 255) / 256), 256, sub_off, n_sub, c->d_sub.p + 2 * n_sub + 5);
            size_t smem = (size_t)TILE * 8 + (size_t)NB2 * 8 + (size_t)NB2 * 4 + ((size_t)NB2 + 2) * 4 + (size_t)TILE * 2 + 16;
            CU(cudaFuncSetAttribute((pc::k_sub_scatter<TILE, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CU(cudaFuncSetAttribute((pc::k_sub_scatter<TILE, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            grid = (unsigned int)std::min<int64_t>(148 * 2, n_tiles);
            c->launches++;
            if (c->scatter_variant == 1) {
                CU(cudaFuncSetAttribute((pc::k_sub_scatter_r<TILE, 512>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                pc::k_sub_scatter_r<TILE, 512><<<grid, 512, smem, c->stream>>>(
                    keys_dev, d_base, d_beg, d_len, c->d_tile.p, (int)n_pairs, n_seg, pb, NB2, sub_off, sub_cur, c->kbuf2.p);
            } else {
                if (2 * k <= PC_PACK_SHIFT && c->scatter_pack == 2)      // level 2: measured no gain (3.83 vs 3.89 ms), off by default
                    pc::k_sub_scatter<TILE, true><<<grid, 256, smem, c->stream>>>(
                        keys_dev, d_base, d_beg, d_len, c->d_tile.p, (int)n_pairs, n_seg, pb, NB2, sub_off, sub_cur, c->kbuf2.p);
                else
                    pc::k_sub_scatter<TILE, false><<<grid, 256, smem, c->stream>>>(
                        keys_dev, d_base, d_beg, d_len, c->d_tile.p, (int)n_pairs, n_seg, pb, NB2, sub_off, sub_cur, c->kbuf2.p);
            }
            CU(cudaGetLastError());
        }
        {
            StageTimer t(c, ST_COUNT);
            if (c->smem_variant == 1) rc = launch_count_smem<4096, 256, 4, false>(c, (unsigned int)n_sub, list_cap, d_nlist, d_novf);
            else if (c->smem_variant == 2) rc = launch_count_smem<16384, 1024, 4, false>(c, (unsigned int)n_sub, list_cap, d_nlist, d_novf);
            else if (c->smem_variant == 3) rc = launch_count_smem<8192, 256, 8, false>(c, (unsigned int)n_sub, list_cap, d_nlist, d_novf);
            else if (c->smem_variant == 4) rc = launch_count_smem<8192, 512, 4, true>(c, (unsigned int)n_sub, list_cap, d_nlist, d_novf);
            else rc = launch_count_smem<8192, 512, 4, false>(c, (unsigned int)n_sub, list_cap, d_nlist, d_novf);
            if (rc) return rc;
        }
    }
    unsigned long long h_nlist = 0, h_maxunit = 0; unsigned int h_novf = 0;
    READ_SMALL(c, &h_nlist, d_nlist, sizeof h_nlist);
    READ_SMALL(c, &h_novf, d_novf, sizeof h_novf);
    READ_SMALL(c, &h_maxunit, c->d_sub.p + 2 * n_sub + 5, sizeof h_maxunit);
    rc = fetch_stats(c); if (rc) return rc;                   // synchronises
    c->max_unit = 0;
    rc = check_unit_size(c, h_maxunit); if (rc) return rc;
    c->n_ovf = h_novf;
    if (h_novf > 0) {
        // sub-buckets whose distinct k-mers did not fit a shared-memory table: count them in one global table and
        // append its survivors to the list
        std::vector<unsigned int> ids(h_novf);
        std::vector<unsigned long long> off(n_sub + 1);
        READ_SMALL(c, ids.data(), c->d_ovf.p, (size_t)h_novf * sizeof(unsigned int));
        READ_SMALL(c, off.data(), sub_off, (size_t)(n_sub + 1) * sizeof(unsigned long long));
        SYNC(c);
        std::sort(ids.begin(), ids.end());
        std::vector<int64_t> fb_begin(h_novf), fb_len(h_novf);
        int64_t fb_total = 0;
        for (unsigned int i = 0; i < h_novf; ++i) {
            fb_begin[i] = (int64_t)off[ids[i]]; fb_len[i] = (int64_t)(off[ids[i] + 1] - off[ids[i]]); fb_total += fb_len[i];
        }
        const uint32_t keep_watch = c->watch_low; c->watch_low = 0;
        rc = count_buckets_table(c, k, 0, 1, c->kbuf2.p, fb_begin.data(), fb_len.data(), (int)h_novf,
                                 (int64_t)(fb_total / c->table_load) + 1024);
        c->watch_low = keep_watch;
        if (rc) return rc;
        unsigned int grid = (unsigned int)std::min<uint64_t>((c->capacity + 1023) / 1024, 148ull * 64);
        LAUNCH(c, pc::k_select, grid, 256, c->table(), c->capacity, c->keep_min, 0xFFFFFFFFu, c->list_keys.p, c->list_counts.p,
               list_cap, d_nlist);
        READ_SMALL(c, &h_nlist, d_nlist, sizeof h_nlist);
        SYNC(c);
        c->pb = pb; c->part_S = 0; c->capacity = (uint64_t)n_sub * (uint64_t)c->smem_slots();
    }
    if (h_nlist > list_cap) return fail(PC_ESTATE, "k-mer list overflow: %llu entries for %llu slots", h_nlist, list_cap);
    c->n_list = (int64_t)h_nlist; c->list_min = c->keep_min;
    c->counted = true; c->count_kind = 1;
    return PC_OK;
}

static bool smem_mode(const pc_ctx* c) { return c->part_mode == 2 || c->part_mode == 3 || c->part_mode < 0; }

// ---------------------------------------------------------------------------------------------- K3, super-k-mer path
// plan[0] = pb (level-1 bucket bits), [1] = nb2 (sub-buckets per level-1 bucket), [2] = m (minimizer length),
// [3] = W = k - m + 1, [4] = nmax (k-mers per record), [5] = 1 when the path can run for this k
struct SkmPlan { int pb; unsigned int nb2; int m, W, nmax; bool ok; };

static SkmPlan skm_plan(const pc_ctx* c, int k, int64_t windows_global, int world) {
    SkmPlan pl{1, 1u, 0, 0, 0, false};
    const int64_t nb_total = std::max<int64_t>(1, windows_global / c->smem_mean());
    // level 1 takes the smaller half of the bits, as in the k-mer path (the short runs of a wide scatter cost more at level 1)
    int pb = 1;
    while (pb < 12 && ((int64_t)2 << (2 * pb)) < nb_total) ++pb;
    if (c->smem_pb > 0) pb = std::min(12, c->smem_pb);
    while ((1 << pb) < world && pb < 12) ++pb;
    const int64_t P = (int64_t)1 << pb;
    pl.pb = pb;
    pl.nb2 = (unsigned int)std::min<int64_t>(PC_SUB_MAX, std::max<int64_t>(1, (nb_total + P - 1) / P));
    pl.m = pc::skm_choose_m(k, P * (int64_t)pl.nb2);
    if (c->skm_m_force > 0 && c->skm_m_force <= k && ((k - c->skm_m_force + 1) & 1) == 0 && k - c->skm_m_force + 1 >= PC_SKM_WMIN &&
        k - c->skm_m_force + 1 <= PC_SKM_WMAX && c->skm_m_force >= 9 && c->skm_m_force <= 16)
        pl.m = c->skm_m_force;                                   // tunable (experiments): minimizer length
    if (pl.m == 0 || k > 32) return pl;
    pl.W = k - pl.m + 1;
    pl.nmax = std::max(1, std::min(std::min(c->skm_nmax, PC_SKM_COUNT_NMAX), PC_SKM_BASES - k + 1));
    pl.ok = true;
    return pl;
}

static void skm_plan_store(const SkmPlan& pl, int32_t plan[8]) {
    memset(plan, 0, 8 * sizeof(int32_t));
    plan[0] = pl.pb; plan[1] = (int32_t)pl.nb2; plan[2] = pl.m; plan[3] = pl.W; plan[4] = pl.nmax; plan[5] = pl.ok ? 1 : 0;
}
static int skm_plan_load(const int32_t plan[8], int k, SkmPlan* pl) {
    if (!plan || !plan[5]) return fail(PC_EINVAL, "super-k-mer plan is not usable for k = %d", k);
    pl->pb = plan[0]; pl->nb2 = (unsigned int)plan[1]; pl->m = plan[2]; pl->W = plan[3]; pl->nmax = plan[4]; pl->ok = true;
    if (pl->pb < 0 || pl->pb > 12 || pl->nb2 < 1 || pl->nb2 > PC_SUB_MAX || pl->m < 1 || pl->m > 16 || pl->W != k - pl->m + 1 ||
        pl->W < PC_SKM_WMIN || pl->W > PC_SKM_WMAX || (pl->W & 1) || pl->nmax < 1 || pl->nmax > PC_SKM_COUNT_NMAX || pl->nmax + k - 1 > PC_SKM_BASES)
        return fail(PC_EINVAL, "bad super-k-mer plan (pb %d nb2 %u m %d W %d nmax %d) for k = %d", pl->pb, pl->nb2, pl->m, pl->W, pl->nmax, k);
    return PC_OK;
}

__global__ void __launch_bounds__(256)
k_skm_hist_split(const unsigned long long* __restrict__ raw, int64_t n, unsigned long long* __restrict__ rec_counts,
                 unsigned long long* __restrict__ kmer_total)
{
    unsigned long long s = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long v = raw[i];
        rec_counts[i] = v & 0xFFFFFFFFull;
        s += v >> 32;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(kmer_total, s);
}

__global__ void __launch_bounds__(256)
k_skm_hist_kmers(const unsigned long long* __restrict__ raw, int64_t n, unsigned long long* __restrict__ kmer_counts)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) kmer_counts[i] = raw[i] >> 32;
}

extern "C++" {
template <int W>
static int launch_skm_extract(pc_ctx* c, int k, const SkmPlan& pl, unsigned int stage_cap) {
    const size_t smem = (size_t)32 * PC_SKM_THREADS * sizeof(uint32_t) + (size_t)stage_cap * sizeof(pc::SkmRec);
    CU(cudaFuncSetAttribute(pc::k_skm_extract<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    c->launches++;
    pc::k_skm_extract<W><<<blocks_for(c->n_words, PC_SKM_THREADS), PC_SKM_THREADS, smem, c->stream>>>(
        c->words.p, c->nmask.p, c->n_words, k, pl.m, pl.nmax, pl.pb, pl.nb2, c->rec_a.p, (unsigned long long)c->skm_cap_records,
        c->d_skm_n.p, c->d_skm_hist.p, stage_cap, c->d_stats.p);
    CU(cudaGetLastError());
    return PC_OK;
}
template <int TILE, int LEVEL>
static int launch_rec_scatter(pc_ctx* c, const pc::SkmRec* recs, const pc::SkmRec* const* d_base, const int64_t* d_beg,
                              const int64_t* d_len, const int64_t* d_tile, int n_pairs, int n_seg, int64_t n_tiles_hint,
                              unsigned int F, unsigned int nb2, const unsigned long long* off, unsigned long long* cursor,
                              pc::SkmRec* out) {
    const size_t smem = (size_t)TILE * 16 + (size_t)F * 8 + (size_t)F * 4 + ((size_t)F + 2) * 4 + (size_t)TILE * 2 + 16;
    CU(cudaFuncSetAttribute((pc::k_rec_scatter<TILE, LEVEL>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = std::max(1, std::min(TILE <= 2048 ? 3 : 2, (int)((size_t)(220 << 10) / (smem + 1024))));
    const unsigned int grid = (unsigned int)std::max<int64_t>(1, std::min<int64_t>(148 * per_sm, n_tiles_hint));
    c->launches++;
    pc::k_rec_scatter<TILE, LEVEL><<<grid, 256, smem, c->stream>>>(recs, d_base, d_beg, d_len, d_tile, n_pairs, n_seg,
                                                                    c->d_skm_n.p, (unsigned long long)c->skm_cap_records, F, nb2,
                                                                    off, cursor, out);
    CU(cudaGetLastError());
    return PC_OK;
}
template <int SLOTS, int THREADS, int GRP>
static int launch_count_skm(pc_ctx* c, const pc::SkmRec* recs, const unsigned long long* sub_off, unsigned int n_sub, int k,
                            unsigned long long list_cap, unsigned long long* d_nlist, unsigned int* d_novf) {
    const size_t smem = (size_t)SLOTS * 12 + (size_t)THREADS * PC_SKM_COUNT_NMAX + (size_t)PC_SKM_NCAND * 2;   // table + one k-mer -> lane map per warp + candidate list
    CU(cudaFuncSetAttribute((pc::k_count_skm<SLOTS, THREADS, GRP>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = std::max(1, (int)((size_t)(220 << 10) / (smem + 1024)));
    per_sm = std::min(per_sm, 2048 / THREADS);
    const unsigned int grid = std::min<unsigned int>(n_sub, 148u * (unsigned int)per_sm);
    c->launches++;
    pc::k_count_skm<SLOTS, THREADS, GRP><<<grid, THREADS, smem, c->stream>>>(
        recs, sub_off, n_sub, k, c->keep_min, c->list_keys.p, c->list_counts.p, list_cap, d_nlist, c->d_ovf.p, d_novf, c->d_stats.p,
        reinterpret_cast<unsigned int*>(c->d_sub.p + 2 * (size_t)n_sub + 7));
    CU(cudaGetLastError());
    return PC_OK;
}
}  // extern "C++"

static int64_t skm_record_cap(const pc_ctx* c, int k, const SkmPlan& pl, int64_t windows) {
    (void)k;
    const double per_window = std::min(1.0, 2.0 / (pl.W + 1) * 1.3 + 1.0 / 32.0 + 1.0 / pl.nmax);
    int64_t cap = (int64_t)((double)windows * per_window) + c->n_words + 65536;
    cap = std::min<int64_t>(cap, windows + 1024);
    return std::max<int64_t>(cap, c->skm_cap_records);           // never shrink (the level-1 buffer may be exported to peers)
}

// extraction + level-1 scatter of the local chunks.  Leaves the bucket-sorted records in rec_b, the raw packed
// histogram ((k-mers << 32) | records per sub-bucket, P * nb2 entries) in d_skm_hist and the level-1 bucket offsets (in
// records) in h_bucket_off.  d_stats must have been reset by the caller.
static int do_skm_partition(pc_ctx* c, int k, const SkmPlan& pl) {
    int rc;
    const int64_t P = (int64_t)1 << pl.pb;
    const int64_t n_sub = P * (int64_t)pl.nb2;
    int64_t windows = 0;
    for (int64_t i = 0; i < c->n_chunks; ++i) windows += std::max<int64_t>(0, c->h_len[i] - k + 1);
    // records: one per minimizer change (~2 / (W + 1) per window in random sequence), one per word boundary inside a
    // run, one per nmax windows of a long run; 30 % head-room.  A genome that needs more (adversarial minimizer order)
    // is reported and the caller falls back to the k-mer path.
    const int64_t cap = skm_record_cap(c, k, pl, windows);
    rc = c->rec_a.alloc((size_t)cap); if (rc) return rc;
    rc = c->rec_b.alloc((size_t)cap); if (rc) return rc;
    c->skm_cap_records = cap;
    rc = c->d_skm_hist.alloc((size_t)n_sub + 8); if (rc) return rc;
    rc = c->d_skm_off.alloc((size_t)n_sub + 8); if (rc) return rc;
    rc = c->d_skm_cur.alloc((size_t)std::max<int64_t>(n_sub, P) + 8); if (rc) return rc;
    rc = c->d_skm_n.alloc(4); if (rc) return rc;
    c->skm_m = pl.m; c->skm_w = pl.W; c->skm_nmax_used = pl.nmax;
    c->stream_valid = false;
    unsigned long long h_n[4] = {0, 0, 0, 0};
    {
        StageTimer t(c, ST_PART_HIST);
        CU(cudaMemsetAsync(c->d_skm_hist.p, 0, ((size_t)n_sub + 8) * sizeof(unsigned long long), c->stream));
        CU(cudaMemsetAsync(c->d_skm_n.p, 0, 4 * sizeof(unsigned long long), c->stream));
        CU(cudaMemsetAsync(c->d_skm_cur.p, 0, ((size_t)std::max<int64_t>(n_sub, P) + 8) * sizeof(unsigned long long), c->stream));
        if (c->n_words > 0) {
            const unsigned int stage_cap = (unsigned int)std::max(256, std::min(c->skm_stage, 8192));
            switch (pl.W) {
#define SKM_CASE(W_) case W_: rc = launch_skm_extract<W_>(c, k, pl, stage_cap); break;
                SKM_CASE(4) SKM_CASE(6) SKM_CASE(8) SKM_CASE(10) SKM_CASE(12) SKM_CASE(14) SKM_CASE(16) SKM_CASE(18) SKM_CASE(20) SKM_CASE(22)
#undef SKM_CASE
                default: return fail(PC_EINVAL, "no extraction kernel for W = %d", pl.W);
            }
            if (rc) return rc;
        }
        // record counts per sub-bucket -> exclusive scan: off[q * nb2 + x]; the level-1 bucket q starts at off[q * nb2]
        LAUNCH(c, k_skm_hist_split, (unsigned int)std::min<int64_t>(148 * 4, (n_sub + 255) / 256), 256, c->d_skm_hist.p, n_sub,
               c->d_skm_off.p, c->d_skm_n.p + 2);
        LAUNCH(c, pc::k_rs_scan<unsigned long long>, 1, 1024, c->d_skm_off.p, n_sub);
    }
    if (c->n_words > 0) {
        StageTimer t(c, ST_PART_SCATTER);
        const int64_t hint = (cap + 2047) / 2048;
        if (P <= 1024) rc = launch_rec_scatter<4096, 1>(c, c->rec_a.p, nullptr, nullptr, nullptr, nullptr, 1, 1, hint, (unsigned int)P, pl.nb2, c->d_skm_off.p, c->d_skm_cur.p, c->rec_b.p);
        else rc = launch_rec_scatter<2048, 1>(c, c->rec_a.p, nullptr, nullptr, nullptr, nullptr, 1, 1, hint, (unsigned int)P, pl.nb2, c->d_skm_off.p, c->d_skm_cur.p, c->rec_b.p);
        if (rc) return rc;
    }
    // bucket offsets (every nb2-th entry of the scan) to the host
    rc = c->d_tile.alloc((size_t)P + 1); if (rc) return rc;
    LAUNCH(c, pc::k_stride_gather, blocks_for(P + 1, 256), 256, reinterpret_cast<const uint64_t*>(c->d_skm_off.p), P + 1, (int64_t)pl.nb2,
           reinterpret_cast<uint64_t*>(c->d_tile.p));
    c->h_bucket_off.assign(P + 1, 0);
    READ_SMALL(c, c->h_bucket_off.data(), c->d_tile.p, (size_t)(P + 1) * sizeof(int64_t));
    READ_SMALL(c, h_n, c->d_skm_n.p, sizeof h_n);
    READ_SMALL(c, &c->h_stats, c->d_stats.p, sizeof(pc::CountStats));
    SYNC(c);
    c->skm_n_records = (int64_t)h_n[0]; c->skm_n_kmers = (int64_t)h_n[2];
    if (h_n[1] != 0 || (int64_t)h_n[0] > cap)
        return fail(PC_ENOMEM, "super-k-mer records overflow their buffer (%llu records for %lld slots)", h_n[0], (long long)cap);
    if ((int64_t)h_n[0] != c->h_bucket_off[P])
        return fail(PC_ESTATE, "super-k-mer histogram (%lld) and record stream (%llu) disagree", (long long)c->h_bucket_off[P], h_n[0]);
    return PC_OK;
}

// level-2 scatter + shared-memory count of n_buckets level-1 buckets given as record segments (one per sender), with
// the packed histogram of exactly those buckets (n_buckets * nb2 entries, summed over the senders) in hist_dev.
static int count_buckets_skm(pc_ctx* c, int k, const SkmPlan& pl, int n_buckets, const pc::SkmRec* recs_dev, const int64_t* seg_begin,
                             const int64_t* seg_len, int n_seg, const pc::SkmRec* const* seg_base_host, const unsigned long long* hist_dev) {
    int rc;
    const unsigned int NB2 = pl.nb2;
    const int64_t n_pairs = (int64_t)n_buckets * n_seg;
    const int64_t n_sub = (int64_t)n_buckets * NB2;
    const int TILE = NB2 <= 1024 ? 4096 : 2048;
    std::vector<int64_t> tile_start(n_pairs + 1, 0);
    int64_t total = 0;
    for (int64_t p = 0; p < n_pairs; ++p) {
        if (seg_len[p] < 0 || seg_begin[p] < 0) return fail(PC_EINVAL, "bad segment");
        total += seg_len[p];
        tile_start[p + 1] = tile_start[p] + (seg_len[p] + TILE - 1) / TILE;
    }
    if (c->keep_min < 1) c->keep_min = 1;
    const int64_t n_tiles = tile_start[n_pairs];
    c->k = k; c->pb = pl.pb; c->sub_nb2 = NB2; c->n_sub = n_sub; c->cross_valid = false;
    c->part_S = 0; c->capacity = (uint64_t)n_sub * (uint64_t)c->smem_slots();

    rc = c->d_seg.alloc((size_t)3 * n_pairs); if (rc) return rc;
    rc = c->d_tile.alloc((size_t)n_pairs + 1); if (rc) return rc;
    rc = c->d_sub.alloc((size_t)2 * n_sub + 8); if (rc) return rc;
    rc = c->d_ovf.alloc((size_t)n_sub + 2); if (rc) return rc;
    rc = c->rec_a.alloc((size_t)std::max<int64_t>(total, 1)); if (rc) return rc;
    rc = c->d_skm_n.alloc(4); if (rc) return rc;
    CU(cudaMemcpyAsync(c->d_seg.p, seg_begin, (size_t)n_pairs * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_seg.p + n_pairs, seg_len, (size_t)n_pairs * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    const int64_t* d_beg = c->d_seg.p;
    const int64_t* d_len = c->d_seg.p + n_pairs;
    const pc::SkmRec* const* d_base = nullptr;
    if (seg_base_host) {
        CU(cudaMemcpyAsync(c->d_seg.p + 2 * n_pairs, seg_base_host, (size_t)n_pairs * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        d_base = reinterpret_cast<const pc::SkmRec* const*>(c->d_seg.p + 2 * n_pairs);
    }
    CU(cudaMemcpyAsync(c->d_tile.p, tile_start.data(), (size_t)(n_pairs + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    unsigned long long* sub_off = c->d_sub.p;                              // [n_sub + 1] record counts, then offsets
    unsigned long long* sub_cur = c->d_sub.p + n_sub + 1;                  // [n_sub] scatter cursors
    unsigned long long* d_nlist = c->d_sub.p + 2 * n_sub + 2;              // [1]
    unsigned int* d_novf = reinterpret_cast<unsigned int*>(c->d_sub.p + 2 * n_sub + 3);
    unsigned long long* d_kmers = c->d_sub.p + 2 * n_sub + 4;              // [1] k-mer instances in these buckets
    CU(cudaMemsetAsync(c->d_sub.p, 0, ((size_t)2 * n_sub + 8) * sizeof(unsigned long long), c->stream));
    unsigned long long h_kmers = 0, h_maxrec = 0;
    {
        StageTimer t(c, ST_PART2);
        LAUNCH(c, k_skm_hist_split, (unsigned int)std::min<int64_t>(148 * 4, (n_sub + 255) / 256), 256, hist_dev, n_sub, sub_off, d_kmers);
        LAUNCH(c, pc::k_rs_scan<unsigned long long>, 1, 1024, sub_off, n_sub);
        LAUNCH(c, pc::k_max_diff, (unsigned int)std::min<int64_t>(148, (n_sub + 255) / 256), 256, sub_off, n_sub, c->d_sub.p + 2 * n_sub + 6);
        READ_SMALL(c, &h_maxrec, c->d_sub.p + 2 * n_sub + 6, sizeof h_maxrec);
        READ_SMALL(c, &h_kmers, d_kmers, sizeof h_kmers);
        if (total > 0) {
            if (TILE == 4096) rc = launch_rec_scatter<4096, 2>(c, recs_dev, d_base, d_beg, d_len, c->d_tile.p, (int)n_pairs, n_seg, n_tiles, NB2, NB2, sub_off, sub_cur, c->rec_a.p);
            else rc = launch_rec_scatter<2048, 2>(c, recs_dev, d_base, d_beg, d_len, c->d_tile.p, (int)n_pairs, n_seg, n_tiles, NB2, NB2, sub_off, sub_cur, c->rec_a.p);
            if (rc) return rc;
        }
    }
    SYNC(c);                                                   // k-mer total: sizes the list
    c->max_unit = 0;
    // (the sub-bucket's own k-mer count travels in 32 bits of the packed histogram: bound it by its records)
    rc = check_unit_size(c, (uint64_t)h_maxrec * (uint64_t)pl.nmax >= (1ull << 32) ? (uint64_t)h_maxrec * (uint64_t)pl.nmax
                                                                                     : std::min<uint64_t>((uint64_t)h_maxrec * (uint64_t)pl.nmax, h_kmers));
    if (rc) return rc;
    const unsigned long long list_cap = h_kmers / c->keep_min + 1024ull;
    rc = c->list_keys.alloc((size_t)list_cap); if (rc) return rc;
    rc = c->list_counts.alloc((size_t)list_cap); if (rc) return rc;
    const bool expand = c->skm_expand != 0;
    if (expand) {
        // records -> 8-byte k-mers in sub-bucket order (kbuf2), then the k-mer path's count kernel and fall-back
        rc = c->d_skm_koff.alloc((size_t)2 * n_sub + 8); if (rc) return rc;
        rc = c->kbuf2.alloc((size_t)std::max<unsigned long long>(h_kmers, 1ull)); if (rc) return rc;
        unsigned long long* koff = c->d_skm_koff.p;                         // [n_sub + 1]
        unsigned long long* kcur = c->d_skm_koff.p + n_sub + 1;             // [n_sub]
        if (total > 0) {
            {
                StageTimer t(c, ST_EXPAND);
                CU(cudaMemsetAsync(kcur, 0, (size_t)n_sub * sizeof(unsigned long long), c->stream));
                LAUNCH(c, k_skm_hist_kmers, (unsigned int)std::min<int64_t>(148 * 4, (n_sub + 255) / 256), 256, hist_dev, n_sub, koff);
                LAUNCH(c, pc::k_rs_scan<unsigned long long>, 1, 1024, koff, n_sub);
                LAUNCH(c, pc::k_rec_expand<512>, (unsigned int)std::min<int64_t>(n_sub, 148 * 4), 512, c->rec_a.p, sub_off, koff, kcur,
                       (unsigned int)n_sub, k, c->kbuf2.p);
            }
            StageTimer t(c, ST_COUNT);
            // k_count_smem reads c->d_sub as the k-mer offsets: hand it the expansion's
            CU(cudaMemcpyAsync(sub_off, koff, ((size_t)n_sub + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
            if (c->smem_variant == 1) rc = launch_count_smem<4096, 256, 4, false>(c, (unsigned int)n_sub, list_cap, d_nlist, d_novf);
            else rc = launch_count_smem<8192, 512, 4, false>(c, (unsigned int)n_sub, list_cap, d_nlist, d_novf);
            if (rc) return rc;
        }
    } else if (total > 0) {
        StageTimer t(c, ST_COUNT);
        if (c->smem_variant == 1) rc = launch_count_skm<4096, 256, 2>(c, c->rec_a.p, sub_off, (unsigned int)n_sub, k, list_cap, d_nlist, d_novf);
        else if (c->smem_variant == 5) rc = launch_count_skm<4096, 512, 2>(c, c->rec_a.p, sub_off, (unsigned int)n_sub, k, list_cap, d_nlist, d_novf);
        else if (c->skm_grp == 4) rc = launch_count_skm<8192, 512, 4>(c, c->rec_a.p, sub_off, (unsigned int)n_sub, k, list_cap, d_nlist, d_novf);
        else rc = launch_count_skm<8192, 512, 2>(c, c->rec_a.p, sub_off, (unsigned int)n_sub, k, list_cap, d_nlist, d_novf);
        if (rc) return rc;
    }
    unsigned long long h_nlist = 0; unsigned int h_novf = 0;
    TRACE("skm count: launched");
    READ_SMALL(c, &h_nlist, d_nlist, sizeof h_nlist);
    READ_SMALL(c, &h_novf, d_novf, sizeof h_novf);
    rc = fetch_stats(c); if (rc) return rc;                   // synchronises
    TRACE("skm count: done, sizes read");
    c->n_ovf = h_novf;
    if (h_novf > 0 && expand) {
        // as in count_buckets_smem: the overflowed sub-buckets are k-mer segments of kbuf2
        std::vector<unsigned int> ids(h_novf);
        std::vector<unsigned long long> off(n_sub + 1);
        READ_SMALL(c, ids.data(), c->d_ovf.p, (size_t)h_novf * sizeof(unsigned int));
        READ_SMALL(c, off.data(), sub_off, (size_t)(n_sub + 1) * sizeof(unsigned long long));
        SYNC(c);
        std::sort(ids.begin(), ids.end());
        std::vector<int64_t> fb_begin(h_novf), fb_len(h_novf);
        int64_t fb_total = 0;
        for (unsigned int i = 0; i < h_novf; ++i) {
            fb_begin[i] = (int64_t)off[ids[i]]; fb_len[i] = (int64_t)(off[ids[i] + 1] - off[ids[i]]); fb_total += fb_len[i];
        }
        const uint32_t keep_watch = c->watch_low; c->watch_low = 0;
        {
            // count_buckets_table times itself as ST_COUNT: keep the shared-memory count's time, report this one apart
            cudaEvent_t k0 = c->ev[ST_COUNT][0], k1 = c->ev[ST_COUNT][1];
            c->ev[ST_COUNT][0] = c->ev[ST_COUNT_FALLBACK][0]; c->ev[ST_COUNT][1] = c->ev[ST_COUNT_FALLBACK][1];
            rc = count_buckets_table(c, k, 0, 1, c->kbuf2.p, fb_begin.data(), fb_len.data(), (int)h_novf,
                                     (int64_t)(fb_total / c->table_load) + 1024);
            c->ev[ST_COUNT_FALLBACK][0] = c->ev[ST_COUNT][0]; c->ev[ST_COUNT_FALLBACK][1] = c->ev[ST_COUNT][1];
            c->ev[ST_COUNT][0] = k0; c->ev[ST_COUNT][1] = k1;
            c->ev_used[ST_COUNT_FALLBACK] = true;
        }
        c->watch_low = keep_watch;
        if (rc) return rc;
        unsigned int grid = (unsigned int)std::min<uint64_t>((c->capacity + 1023) / 1024, 148ull * 64);
        LAUNCH(c, pc::k_select, grid, 256, c->table(), c->capacity, c->keep_min, 0xFFFFFFFFu, c->list_keys.p, c->list_counts.p,
               list_cap, d_nlist);
        READ_SMALL(c, &h_nlist, d_nlist, sizeof h_nlist);
        SYNC(c);
        c->pb = pl.pb; c->part_S = 0; c->capacity = (uint64_t)n_sub * (uint64_t)c->smem_slots();
    } else if (h_novf > 0) {
        // sub-buckets whose distinct k-mers did not fit a shared-memory table: one global table for all of them
        LAUNCH(c, pc::k_hist_kmer_total, std::min<unsigned int>((h_novf + 255u) / 256u, 148u), 256, hist_dev, c->d_ovf.p, h_novf, d_kmers + 1);
        unsigned long long fb_total = 0;
        READ_SMALL(c, &fb_total, d_kmers + 1, sizeof fb_total);
        SYNC(c);
        TRACE("skm fall-back: size read");
        rc = table_prepare(c, k, (int64_t)((double)fb_total / c->table_load) + 1024, /*reset_stats=*/false); if (rc) return rc;
        TRACE("skm fall-back: table ready");
        {
            StageTimer t(c, ST_COUNT_FALLBACK);
            // a heavy minimizer (one m-mer inside a high-copy repeat family carries the mutated k-mers of every copy) can put
            // millions of records into one sub-bucket: 32 CTAs per listed sub-bucket
            dim3 fgrid(std::min<unsigned int>(h_novf, 65535u), 32u);
            c->launches++;
            pc::k_count_rec_table<<<fgrid, 256, 0, c->stream>>>(c->rec_a.p, sub_off, c->d_ovf.p, h_novf, k, c->table(), c->capacity, c->d_stats.p);
            CU(cudaGetLastError());
        }
        rc = fetch_stats(c); if (rc) return rc;
        unsigned int grid = (unsigned int)std::min<uint64_t>((c->capacity + 1023) / 1024, 148ull * 64);
        LAUNCH(c, pc::k_select, grid, 256, c->table(), c->capacity, c->keep_min, 0xFFFFFFFFu, c->list_keys.p, c->list_counts.p,
               list_cap, d_nlist);
        READ_SMALL(c, &h_nlist, d_nlist, sizeof h_nlist);
        SYNC(c);
        c->part_S = 0; c->capacity = (uint64_t)n_sub * (uint64_t)c->smem_slots();
    }
    if (h_nlist > list_cap) return fail(PC_ESTATE, "k-mer list overflow: %llu entries for %llu slots", h_nlist, list_cap);
    c->n_list = (int64_t)h_nlist; c->list_min = c->keep_min;
    c->skm_n_kmers = (int64_t)h_kmers;
    c->counted = true; c->count_kind = 1; c->skm_used = true;
    return PC_OK;
}

static int do_count_buckets(pc_ctx* c, int k, int pb, int n_buckets, const uint64_t* keys_dev, const int64_t* seg_begin,
                            const int64_t* seg_len, int n_seg, int64_t capacity, const uint64_t* const* seg_base_host = nullptr) {
    if (smem_mode(c)) return count_buckets_smem(c, k, pb, n_buckets, keys_dev, seg_begin, seg_len, n_seg, seg_base_host);
    return count_buckets_table(c, k, pb, n_buckets, keys_dev, seg_begin, seg_len, n_seg, capacity, seg_base_host);
}

// Free device buffers that later stages no longer need (large genomes on one GPU: the count table, the bucket buffer
// and the hit buffers are each tens of GB at Gbp scale and are never needed at the same time).
int pc_release(pc_ctx* c, int what) {
    int rc = use_device(c); if (rc) return rc;
    SYNC(c);
    if (c->zeroed_slots > 0) { CU(cudaStreamSynchronize(c->aux_stream)); c->zeroed_slots = 0; }
    if (what & PC_REL_ASCII) { c->ascii.release(); c->ascii_padded = 0; }
    if (what & PC_REL_BUCKETS) { c->kbuf.release(); c->kbuf2.release(); c->kbuf_shared = false; c->rec_a.release(); c->rec_b.release(); c->skm_cap_records = 0; c->rec_shared = false; }
    if (what & PC_REL_TABLE) {
        c->table_keys.release(); c->table_counts.release(); c->counted = false; c->cross_valid = false; c->cross_keys.release();
        c->list_keys.release(); c->list_counts.release(); c->n_list = 0;
    }
    if (what & PC_REL_STREAM) { c->kstream.release(); c->stream_valid = false; }
    if (what & PC_REL_HITS) { c->hits_a.release(); c->hits_b.release(); }
    return PC_OK;
}

int pc_watch_threshold(pc_ctx* c, uint32_t low) {
    if (!c) return fail(PC_EINVAL, "null context");
    c->watch_low = low;
    return PC_OK;
}

int pc_tune(pc_ctx* c, const char* key, int64_t value) {
    int rc = use_device(c); if (rc) return rc;
    if (!key) return fail(PC_EINVAL, "null key");
    std::string k(key);
    if (k == "count_batch") c->count_batch = (int)value;
    else if (k == "count_minb") c->count_minb = (int)value;
    else if (k == "count_mode") c->count_mode = (int)value;
    else if (k == "part_mode") c->part_mode = (int)value;
    else if (k == "count_items") c->count_items = (int)value;
    else if (k == "probe_mode") c->probe_mode = (int)value;
    else if (k == "hash_mode") {
        int m = (int)value;
        if (m < 0 || m > 1) return fail(PC_EINVAL, "hash_mode must be 0 or 1");
        CU(cudaMemcpyToSymbol(pc::g_hash_mode, &m, sizeof m));
        c->counted = c->selected = c->built = false;          // tables built with the other hash are unusable
        c->stream_valid = false;
    }
    else if (k == "keep_min") c->keep_min = (uint32_t)std::max<int64_t>(1, value);
    else if (k == "smem_target") c->smem_target = value <= 0 ? 0 : std::max<int64_t>(64, value);
    else if (k == "smem_variant") c->smem_variant = (int)value;
    else if (k == "scatter_variant") c->scatter_variant = (int)value;
    else if (k == "smem_pb") c->smem_pb = (int)value;
    else if (k == "skm_nmax") c->skm_nmax = (int)std::max<int64_t>(1, std::min<int64_t>(PC_SKM_COUNT_NMAX, value));
    else if (k == "skm_min_k") c->skm_min_k = (int)value;
    else if (k == "skm_grp") c->skm_grp = (int)value;
    else if (k == "skm_m") c->skm_m_force = (int)value;
    else if (k == "skm_expand") c->skm_expand = (int)value;
    else if (k == "skm_stage") c->skm_stage = (int)value;
    else if (k == "scatter_tile") c->scatter_tile = (int)value;
    else if (k == "scatter_pack") c->scatter_pack = (int)value;
    else if (k == "rowsort_warps") c->rowsort_warps = (int)value;
    else if (k == "select_raw") c->select_raw = (int)value;
    else if (k == "gram_variant") c->gram_variant = (int)value;
    else if (k == "gram_preround") c->gram_preround = value ? 1 : 0;
    else if (k == "csr_mode") c->csr_mode = (int)value;
    else if (k == "csr_range_bits") c->csr_range_bits = value <= 0 ? 0 : ((std::max<int64_t>(1024, value) + 1023) / 1024) * 1024;
    else if (k == "csr_ncap") c->csr_ncap = std::max<int64_t>(0, value);
    else if (k == "prezero") c->prezero = (int)value;
    else if (k == "probe_fp") c->probe_fp = (int)value;
    else if (k == "probe_filter") c->probe_filter = (int)value;
    else if (k == "probe_threads") c->probe_threads = (int)value;
    else if (k == "filter_bits") c->filter_bits = std::max<int64_t>(1024, value);
    else if (k == "part_casfirst") c->part_casfirst = (int)value;
    else if (k == "part_minb") c->part_minb = (int)value;
    else if (k == "probe_batch") c->probe_batch = (int)value;
    else if (k == "probe_roll") c->probe_roll = (int)value;
    else if (k == "probe_table8") c->probe_table8 = (int)value;
    else if (k == "rep_load_inv") c->rep_load_inv = (int)value;
    else if (k == "part_mbytes") c->part_bytes = std::max<int64_t>(1, value) << 20;
    else if (k == "part_blocks") c->part_blocks = (int)std::max<int64_t>(0, value);
    else if (k == "table_load_pct") c->table_load = std::min(0.95, std::max(0.05, value / 100.0));
    else if (k == "l2_fetch_granularity") CU(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value));
    else return fail(PC_EINVAL, "unknown tunable %s", key);
    return PC_OK;
}

int pc_count(pc_ctx* c, int k, int64_t capacity) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->loaded) return fail(PC_ESTATE, "pc_count before pc_load_sequences");
    if (k < 1 || k > 32) return fail(PC_EINVAL, "k = %d unsupported: this path handles 1 <= k <= 32 (uint64 k-mers)", k);
    c->counted = c->selected = c->built = c->tfidf_done = false;
    c->stream_valid = false;
    ReliefScope relief(c);
    int64_t windows = 0;
    for (int64_t i = 0; i < c->n_chunks; ++i) windows += std::max<int64_t>(0, c->h_len[i] - k + 1);
    if (capacity <= 0) {
        // distinct canonical k-mers <= windows and <= 4^k (palindrome-aware bound not needed: it is an upper bound)
        double space = k >= 31 ? 9.3e18 : (double)(1ull << (2 * k));
        double bound = std::min((double)windows, space);
        capacity = (int64_t)(bound / c->table_load) + 1024;
    }
    // partition plan: P = 2^pb sub-tables of S slots, each about part_bytes so that it stays L2 resident
    int pb = 0;
    bool partitioned = c->part_mode == 1 || c->part_mode == 2 || c->part_mode == 3 ||
                       (c->part_mode < 0 && (size_t)capacity * (size_t)PC_SLOT_BYTES > (size_t)(64ll << 20));
    c->skm_used = false;
    if (partitioned && c->n_words > 0 && (c->part_mode == 3 || (c->part_mode < 0 && k >= c->skm_min_k))) {
        // super-k-mer path: records instead of k-mers through the two scatters, expansion inside the count kernel
        SkmPlan pl = skm_plan(c, k, windows, 1);
        if (!pl.ok && c->part_mode == 3)
            return fail(PC_EINVAL, "k = %d is too short for the super-k-mer count (part_mode 3): use part_mode 2", k);
        if (pl.ok) {
            CU(cudaMemsetAsync(c->d_stats.p, 0, sizeof(pc::CountStats), c->stream));
            c->relief_path = 2;
            rc = do_skm_partition(c, k, pl);
            if (rc == PC_OK) {
                const int64_t Pn = (int64_t)1 << pl.pb;
                std::vector<int64_t> seg_begin(Pn), seg_len(Pn);
                for (int64_t q = 0; q < Pn; ++q) { seg_begin[q] = c->h_bucket_off[q]; seg_len[q] = c->h_bucket_off[q + 1] - c->h_bucket_off[q]; }
                return count_buckets_skm(c, k, pl, (int)Pn, c->rec_b.p, seg_begin.data(), seg_len.data(), 1, nullptr, c->d_skm_hist.p);
            }
            if (rc != PC_ENOMEM || c->part_mode == 3) return rc;
            // more records than the buffer plans for (adversarial minimizer order): the k-mer path below always works
        }
    }
    if (partitioned && smem_mode(c)) {
        // two-level split into about windows / smem_target sub-buckets: level 1 takes the square root, rounded to a
        // power of two, the level-2 fan-out is fixed once the exact number of valid windows is known
        int64_t nb_total = std::max<int64_t>(1, windows / c->smem_mean());
        // (measured at 4.46e9 k-mers: level 1 / level 2 cost 23 + 55 ms at pb 9, 28 + 44 at 10, 44 + 37 at 11, 86 + 36 at
        // 12 -- the level-1 scatter suffers more from short runs than level 2, so level 1 takes the smaller half)
        pb = 1;
        while (pb < 12 && ((int64_t)2 << (2 * pb)) < nb_total) ++pb;
        if (c->smem_pb > 0) pb = std::min(12, c->smem_pb);       // tunable override
    } else if (partitioned) {
        while (pb < 12 && ((size_t)capacity * (size_t)PC_SLOT_BYTES) >> pb > (size_t)c->part_bytes) ++pb;
        if (pb == 0) pb = 1;
    }
    uint64_t P = 1ull << pb;
    c->max_unit = 0;
    if (!partitioned || c->n_words == 0) {
        rc = check_unit_size(c, (uint64_t)windows); if (rc) return rc;
        uint64_t S = std::max<uint64_t>(64, (uint64_t)capacity);
        rc = table_prepare(c, k, (int64_t)S); if (rc) return rc;
        c->pb = 0; c->part_S = S;
        if (c->n_words > 0) {
            StageTimer t(c, ST_COUNT);
            rc = launch_count(c); if (rc) return rc;
        }
        rc = fetch_stats(c); if (rc) return rc;
        c->counted = true; c->count_kind = 0;
        return PC_OK;
    }

    // ---- partitioned path: extract -> tile scatter -> bucket-by-bucket count ----------------------------------
    c->relief_path = 1;
    CU(cudaMemsetAsync(c->d_stats.p, 0, sizeof(pc::CountStats), c->stream));
    if (!smem_mode(c)) { rc = table_prezero(c, k, (int64_t)(((uint64_t)capacity + P - 1) / P * P)); if (rc) return rc; }
    rc = do_partition(c, k, pb); if (rc) return rc;
    std::vector<int64_t> seg_begin(P), seg_len(P);
    for (uint64_t q = 0; q < P; ++q) { seg_begin[q] = c->h_bucket_off[q]; seg_len[q] = c->h_bucket_off[q + 1] - c->h_bucket_off[q]; }
    rc = do_count_buckets(c, k, pb, (int)P, c->kbuf.p, seg_begin.data(), seg_len.data(), 1, capacity); if (rc) return rc;
    return PC_OK;
}

// Multi-GPU building blocks: the bucket id is the top pb bits of the k-mer hash and rank r owns a
// contiguous range of buckets, so after pc_partition the k-mers every peer must receive are one contiguous slice of the
// bucket-sorted buffer: they travel as raw 8-byte k-mers in one all-to-all over NVLink and are counted once, by their
// owner, with the same L2-resident bucket kernel the single-GPU path uses.
int pc_partition(pc_ctx* c, int k, int pb, int64_t* bucket_off, uint64_t** keys_dev) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->loaded) return fail(PC_ESTATE, "pc_partition before pc_load_sequences");
    if (k < 1 || k > 32) return fail(PC_EINVAL, "k = %d unsupported: this path handles 1 <= k <= 32 (uint64 k-mers)", k);
    if (pb < 0 || pb > 12 || !bucket_off || !keys_dev) return fail(PC_EINVAL, "bad pc_partition arguments");
    c->counted = c->selected = c->built = c->tfidf_done = false;
    c->stream_valid = false;
    CU(cudaMemsetAsync(c->d_stats.p, 0, sizeof(pc::CountStats), c->stream));
    {   // the owner will count about as many k-mers as this rank extracts: start zeroing a table of that size now
        int64_t windows = 0;
        for (int64_t i = 0; i < c->n_chunks; ++i) windows += std::max<int64_t>(0, c->h_len[i] - k + 1);
        if (!smem_mode(c)) { rc = table_prezero(c, k, (int64_t)(windows / c->table_load) + 1024); if (rc) return rc; }
    }
    rc = do_partition(c, k, pb); if (rc) return rc;
    memcpy(bucket_off, c->h_bucket_off.data(), ((size_t)(1u << pb) + 1) * sizeof(int64_t));
    *keys_dev = c->kbuf.p;
    return PC_OK;
}

// ---- peer-memory variant of the exchange: owners read the senders' bucket buffers directly over NVLink ------------
int pc_kbuf_export(pc_ctx* c, int64_t n_keys, void* ipc_handle) {
    int rc = use_device(c); if (rc) return rc;
    if (n_keys < 1 || !ipc_handle) return fail(PC_EINVAL, "bad pc_kbuf_export arguments");
    rc = c->kbuf.alloc((size_t)n_keys); if (rc) return rc;           // kept (and re-used by pc_partition) while it is big enough
    c->kbuf_shared = true;
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, c->kbuf.p));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(ipc_handle, &h, sizeof h);
    return PC_OK;
}

int pc_peer_open(pc_ctx* c, const void* ipc_handle, uint64_t** dev_ptr) {
    int rc = use_device(c); if (rc) return rc;
    if (!ipc_handle || !dev_ptr) return fail(PC_EINVAL, "bad pc_peer_open arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof h);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr = static_cast<uint64_t*>(p);
    return PC_OK;
}

int pc_peer_close(pc_ctx* c, uint64_t* dev_ptr) {
    int rc = use_device(c); if (rc) return rc;
    if (dev_ptr) CU(cudaIpcCloseMemHandle(dev_ptr));
    return PC_OK;
}

int pc_count_buckets_from(pc_ctx* c, int k, int pb, int n_buckets, const uint64_t* const* seg_base, const int64_t* seg_begin,
                          const int64_t* seg_len, int n_seg, int64_t capacity) {
    int rc = use_device(c); if (rc) return rc;
    if (k < 1 || k > 32) return fail(PC_EINVAL, "k = %d unsupported: this path handles 1 <= k <= 32 (uint64 k-mers)", k);
    if (pb < 0 || pb > 12 || n_buckets < 1 || n_buckets > (1 << pb) || n_seg < 1 || !seg_base || !seg_begin || !seg_len)
        return fail(PC_EINVAL, "bad pc_count_buckets_from arguments");
    c->counted = c->selected = c->built = c->tfidf_done = false;
    unsigned long long keep = c->h_stats.valid_windows;
    pc::CountStats z{}; z.valid_windows = keep;
    CU(cudaMemcpyAsync(c->d_stats.p, &z, sizeof z, cudaMemcpyHostToDevice, c->stream));
    SYNC(c);
    return do_count_buckets(c, k, pb, n_buckets, nullptr, seg_begin, seg_len, n_seg, capacity, seg_base);
}

int pc_count_buckets(pc_ctx* c, int k, int pb, int n_buckets, const uint64_t* keys_dev, const int64_t* seg_begin,
                     const int64_t* seg_len, int n_seg, int64_t capacity) {
    int rc = use_device(c); if (rc) return rc;
    if (k < 1 || k > 32) return fail(PC_EINVAL, "k = %d unsupported: this path handles 1 <= k <= 32 (uint64 k-mers)", k);
    if (pb < 0 || pb > 12 || n_buckets < 1 || n_buckets > (1 << pb) || n_seg < 1 || !seg_begin || !seg_len)
        return fail(PC_EINVAL, "bad pc_count_buckets arguments");
    c->counted = c->selected = c->built = c->tfidf_done = false;
    unsigned long long keep = c->h_stats.valid_windows;        // pc_partition already counted the local windows
    pc::CountStats z{}; z.valid_windows = keep;
    CU(cudaMemcpyAsync(c->d_stats.p, &z, sizeof z, cudaMemcpyHostToDevice, c->stream));
    SYNC(c);
    return do_count_buckets(c, k, pb, n_buckets, keys_dev, seg_begin, seg_len, n_seg, capacity);
}

// ---- sender-side level-2 histogram (multi-GPU): the owner would otherwise read every peer segment twice over NVLink
int pc_sub_plan(pc_ctx* c, int64_t total, int n_buckets) {
    if (!c || total < 0 || n_buckets < 1) return 0;
    if (!smem_mode(c)) return 0;
    return (int)sub_plan(c, total, n_buckets);
}

int pc_sub_hist(pc_ctx* c, int pb, int nb2, uint64_t* hist_dev) {
    int rc = use_device(c); if (rc) return rc;
    if (pb < 0 || pb > 12 || nb2 < 1 || nb2 > PC_SUB_MAX || !hist_dev) return fail(PC_EINVAL, "bad pc_sub_hist arguments");
    const int64_t P = (int64_t)1 << pb;
    if ((int64_t)c->h_bucket_off.size() != P + 1) return fail(PC_ESTATE, "pc_sub_hist needs the buckets of pc_partition(pb = %d)", pb);
    constexpr int TILE = 8192;
    std::vector<int64_t> desc(3 * P + 1, 0);                   // seg_begin[P], seg_len[P], tile_start[P + 1]
    for (int64_t q = 0; q < P; ++q) {
        desc[q] = c->h_bucket_off[q]; desc[P + q] = c->h_bucket_off[q + 1] - c->h_bucket_off[q];
        desc[2 * P + q + 1] = desc[2 * P + q] + (desc[P + q] + TILE - 1) / TILE;
    }
    const int64_t n_tiles = desc[3 * P];
    rc = c->d_tile.alloc((size_t)3 * P + 1); if (rc) return rc;
    CU(cudaMemcpyAsync(c->d_tile.p, desc.data(), desc.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(hist_dev, 0, (size_t)P * nb2 * sizeof(uint64_t), c->stream));
    if (n_tiles > 0) {
        StageTimer t(c, ST_PART2);
        unsigned int grid = (unsigned int)std::min<int64_t>(148 * 8, n_tiles);
        c->launches++;
        pc::k_sub_hist<TILE><<<grid, 256, (size_t)nb2 * sizeof(unsigned int), c->stream>>>(
            c->kbuf.p, nullptr, c->d_tile.p, c->d_tile.p + P, c->d_tile.p + 2 * P, (int)P, 1, pb, (unsigned int)nb2,
            reinterpret_cast<unsigned long long*>(hist_dev));
        CU(cudaGetLastError());
    }
    SYNC(c);                      // the host descriptor goes away
    return PC_OK;
}

int pc_sub_hist_set(pc_ctx* c, int nb2, const uint64_t* hist_dev, int64_t n) {
    if (!c) return fail(PC_EINVAL, "null context");
    if (!hist_dev || nb2 < 1 || nb2 > PC_SUB_MAX || n < nb2 || n % nb2) return fail(PC_EINVAL, "bad pc_sub_hist_set arguments");
    c->staged_hist = reinterpret_cast<const unsigned long long*>(hist_dev); c->staged_n = n; c->staged_nb2 = (unsigned int)nb2;
    return PC_OK;
}

// ---- super-k-mer building blocks for several GPUs (same roles as pc_partition / pc_kbuf_export / pc_count_buckets_from)
int pc_skm_plan(pc_ctx* c, int k, int64_t windows_global, int world, int32_t plan[8]) {
    if (!c || !plan || k < 1 || k > 32 || windows_global < 0 || world < 1) return fail(PC_EINVAL, "bad pc_skm_plan arguments");
    SkmPlan pl = skm_plan(c, k, windows_global, world);
    if (!(c->part_mode == 3 || (c->part_mode < 0 && k >= c->skm_min_k))) pl.ok = false;      // the rule of pc_count
    skm_plan_store(pl, plan);
    return PC_OK;
}

int pc_skm_partition(pc_ctx* c, int k, const int32_t plan[8], int64_t* bucket_off, uint64_t** recs_dev, uint64_t** hist_dev) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->loaded) return fail(PC_ESTATE, "pc_skm_partition before pc_load_sequences");
    if (k < 1 || k > 32 || !bucket_off || !recs_dev || !hist_dev) return fail(PC_EINVAL, "bad pc_skm_partition arguments");
    SkmPlan pl; rc = skm_plan_load(plan, k, &pl); if (rc) return rc;
    c->counted = c->selected = c->built = c->tfidf_done = false;
    CU(cudaMemsetAsync(c->d_stats.p, 0, sizeof(pc::CountStats), c->stream));
    rc = do_skm_partition(c, k, pl); if (rc) return rc;
    memcpy(bucket_off, c->h_bucket_off.data(), ((size_t)(1u << pl.pb) + 1) * sizeof(int64_t));
    *recs_dev = reinterpret_cast<uint64_t*>(c->rec_b.p);
    *hist_dev = reinterpret_cast<uint64_t*>(c->d_skm_hist.p);
    return PC_OK;
}

int pc_skm_export(pc_ctx* c, int k, const int32_t plan[8], void* ipc_handle) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->loaded) return fail(PC_ESTATE, "pc_skm_export before pc_load_sequences");
    if (!ipc_handle) return fail(PC_EINVAL, "bad pc_skm_export arguments");
    SkmPlan pl; rc = skm_plan_load(plan, k, &pl); if (rc) return rc;
    int64_t windows = 0;
    for (int64_t i = 0; i < c->n_chunks; ++i) windows += std::max<int64_t>(0, c->h_len[i] - k + 1);
    const int64_t cap = skm_record_cap(c, k, pl, windows);       // what pc_skm_partition will need: kept while it is big enough
    rc = c->rec_b.alloc((size_t)cap); if (rc) return rc;
    rc = c->rec_a.alloc((size_t)cap); if (rc) return rc;
    c->skm_cap_records = cap;
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, c->rec_b.p));
    memcpy(ipc_handle, &h, sizeof h);
    c->rec_shared = true;
    return PC_OK;
}

int pc_skm_count_from(pc_ctx* c, int k, const int32_t plan[8], int n_buckets, const uint64_t* recs_dev,
                      const uint64_t* const* seg_base, const int64_t* seg_begin, const int64_t* seg_len, int n_seg,
                      const uint64_t* hist_dev) {
    int rc = use_device(c); if (rc) return rc;
    SkmPlan pl; rc = skm_plan_load(plan, k, &pl); if (rc) return rc;
    if (n_buckets < 1 || n_buckets > (1 << pl.pb) || n_seg < 1 || (!seg_base && !recs_dev) || !seg_begin || !seg_len || !hist_dev)
        return fail(PC_EINVAL, "bad pc_skm_count_from arguments");
    c->counted = c->selected = c->built = c->tfidf_done = false;
    unsigned long long keep = c->h_stats.valid_windows;        // pc_skm_partition already counted the local windows
    pc::CountStats z{}; z.valid_windows = keep;
    CU(cudaMemcpyAsync(c->d_stats.p, &z, sizeof z, cudaMemcpyHostToDevice, c->stream));
    SYNC(c);
    return count_buckets_skm(c, k, pl, n_buckets, reinterpret_cast<const pc::SkmRec*>(recs_dev), seg_begin, seg_len, n_seg,
                             reinterpret_cast<const pc::SkmRec* const*>(seg_base), reinterpret_cast<const unsigned long long*>(hist_dev));
}

// ---- N4: externally counted k-mers (a .kcount from HipMer / jellyfish / kmercountexact, format.py:123-133) become the
// count result, so that K4..K7 run on them unchanged
int pc_load_counts(pc_ctx* c, const uint64_t* keys, const uint32_t* counts, int64_t n, int k, int loc) {
    int rc = use_device(c); if (rc) return rc;
    if (k < 1 || k > 32) return fail(PC_EINVAL, "k = %d unsupported: this path handles 1 <= k <= 32 (uint64 k-mers)", k);
    if (n < 0 || (n > 0 && (!keys || !counts))) return fail(PC_EINVAL, "bad pc_load_counts arguments");
    c->counted = c->selected = c->built = c->tfidf_done = false;
    c->skm_used = false; c->cross_valid = false;
    rc = c->list_keys.alloc((size_t)std::max<int64_t>(n, 1)); if (rc) return rc;
    rc = c->list_counts.alloc((size_t)std::max<int64_t>(n, 1)); if (rc) return rc;
    rc = c->sel_a.alloc(std::max<int64_t>(n, 1)); if (rc) return rc;
    rc = c->sel_b.alloc(std::max<int64_t>(n, 1)); if (rc) return rc;
    const cudaMemcpyKind kind = loc == PC_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    unsigned long long bad[2] = {0, 0};
    if (n > 0) {
        CU(cudaMemcpyAsync(c->list_keys.p, keys, (size_t)n * sizeof(uint64_t), kind, c->stream));
        CU(cudaMemcpyAsync(c->list_counts.p, counts, (size_t)n * sizeof(uint32_t), kind, c->stream));
        LAUNCH(c, pc::k_canonicalize, blocks_for(n, 256), 256, c->list_keys.p, n, k);
        // every k-mer once: a list that names a k-mer twice (or both of its strands) would double its matrix column
        CU(cudaMemcpyAsync(c->sel_a.p, c->list_keys.p, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c->stream));
        uint64_t* sorted = nullptr;
        rc = radix_sort_u64(c, c->sel_a.p, c->sel_b.p, n, 2 * k, &sorted); if (rc) return rc;
        CU(cudaMemsetAsync(c->d_counter.p, 0, 2 * sizeof(unsigned long long), c->stream));
        LAUNCH(c, pc::k_count_dups, blocks_for(n, 256), 256, sorted, n, c->d_counter.p);
        LAUNCH(c, pc::k_count_sentinel, blocks_for(n, 256), 256, c->list_keys.p, n, c->d_counter.p + 1);
        READ_SMALL(c, bad, c->d_counter.p, sizeof bad);
    }
    SYNC(c);
    if (bad[1]) return fail(PC_EINVAL, "%llu entries are not k-mers of length %d (N or other letters)", bad[1], k);
    if (bad[0]) return fail(PC_EINVAL, "%llu k-mers are listed more than once after strand collapse: merge their counts first", bad[0]);
    c->k = k; c->n_list = n; c->list_min = 1; c->keep_min = 1;
    c->h_stats = pc::CountStats{}; c->h_stats.distinct = (unsigned long long)n;
    c->capacity = (uint64_t)std::max<int64_t>(n, 1); c->pb = 0; c->n_sub = 0; c->sub_nb2 = 0; c->n_ovf = 0;
    c->counted = true; c->count_kind = 1;
    return PC_OK;
}

int pc_count_info(pc_ctx* c, int64_t info[8]) {
    if (!c || !c->counted) return fail(PC_ESTATE, "no count table");
    memset(info, 0, 8 * sizeof(int64_t));
    info[0] = (int64_t)c->capacity; info[1] = (int64_t)c->h_stats.distinct; info[2] = (int64_t)c->h_stats.valid_windows;
    info[3] = (int64_t)(c->capacity * (size_t)PC_SLOT_BYTES); info[4] = (int64_t)c->h_stats.max_probe; info[5] = c->k;
    info[6] = c->count_kind == 1 ? c->n_list : 0; info[7] = c->pb;
    if (c->count_kind == 1) info[3] = (int64_t)((size_t)c->list_keys.n * 12 + (c->skm_used ? (size_t)c->rec_a.n * 16 : (size_t)c->kbuf2.n * 8));
    return PC_OK;
}

int pc_count_flags(pc_ctx* c, int64_t flags[4]) {
    if (!c || !c->counted || !flags) return fail(PC_ESTATE, "no count result");
    flags[0] = c->max_unit >= (1ull << 32) ? 1 : 0;            // always 0 for a count that was returned: see check_unit_size
    flags[1] = (int64_t)c->max_unit;
    flags[2] = c->count_kind == 1 ? c->n_ovf : 0;
    flags[3] = c->count_kind == 1 && c->n_sub > 0 ? (int64_t)(c->capacity / (uint64_t)c->n_sub) : (int64_t)c->capacity;
    return PC_OK;
}

int pc_count_detail(pc_ctx* c, int64_t detail[8]) {
    if (!c || !c->counted) return fail(PC_ESTATE, "no count result");
    memset(detail, 0, 8 * sizeof(int64_t));
    detail[0] = c->count_kind;
    if (c->count_kind == 1) {
        detail[1] = c->n_list; detail[2] = c->list_min; detail[3] = c->n_sub; detail[4] = c->sub_nb2; detail[5] = c->n_ovf;
        if (c->skm_used) { detail[6] = c->skm_n_records; detail[7] = c->skm_m; }
    }
    return PC_OK;
}

// runs k_select into (keys_dev, counts_dev); returns the number of survivors
static int select_into(pc_ctx* c, uint32_t low, uint32_t high, uint64_t* keys_dev, uint32_t* counts_dev, int64_t cap,
                       int64_t* n_out) {
    CU(cudaMemsetAsync(c->d_counter.p, 0, sizeof(unsigned long long), c->stream));
    if (c->count_kind == 1) {
        if (low < c->list_min)
            return fail(PC_ESTATE, "k-mers with fewer than %u occurrences were not kept by this count (pc_tune keep_min); asked for >= %u",
                        c->list_min, low);
        if (c->n_list > 0) {
            unsigned int grid = (unsigned int)std::min<int64_t>((c->n_list + 1023) / 1024, 148ll * 16);
            LAUNCH(c, pc::k_select_list, grid, 256, c->list_keys.p, c->list_counts.p, c->n_list, low, high, keys_dev, counts_dev,
                   (unsigned long long)cap, c->d_counter.p);
        }
    } else {
    unsigned int grid = (unsigned int)std::min<uint64_t>((c->capacity + 1023) / 1024, 148ull * 64);
    LAUNCH(c, pc::k_select, grid, 256, c->table(), c->capacity, low, high, keys_dev, counts_dev,
           (unsigned long long)cap, c->d_counter.p);
    }
    unsigned long long n = 0;
    READ_SMALL(c, &n, c->d_counter.p, sizeof n);
    SYNC(c);
    *n_out = (int64_t)n;
    return PC_OK;
}

int pc_dump_counts(pc_ctx* c, uint32_t min_count, uint32_t max_count, uint64_t* keys, uint32_t* counts, int64_t cap,
                   int loc, int64_t* n) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->counted) return fail(PC_ESTATE, "pc_dump_counts before pc_count");
    if (!n) return fail(PC_EINVAL, "n is null");
    if (!keys) return select_into(c, min_count, max_count, nullptr, nullptr, 0, n);
    if (loc == PC_DEVICE) {
        rc = select_into(c, min_count, max_count, keys, counts, cap, n); if (rc) return rc;
        if (*n > cap) return fail(PC_EINVAL, "dump buffer too small: need %lld", (long long)*n);
        return PC_OK;
    }
    DevBuf<uint64_t> dk; DevBuf<uint32_t> dc;
    rc = dk.alloc(cap); if (rc) return rc;
    rc = dc.alloc(cap); if (rc) { dk.release(); return rc; }
    rc = select_into(c, min_count, max_count, dk.p, dc.p, cap, n);
    if (!rc && *n > cap) rc = fail(PC_EINVAL, "dump buffer too small: need %lld", (long long)*n);
    if (!rc) rc = copy_out(c, keys, dk.p, *n * sizeof(uint64_t), PC_HOST);
    if (!rc && counts) rc = copy_out(c, counts, dc.p, *n * sizeof(uint32_t), PC_HOST);
    dk.release(); dc.release();
    return rc;
}

int pc_export_by_owner(pc_ctx* c, int world, uint64_t* keys_dev, uint32_t* counts_dev, int64_t cap, int64_t* offsets) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->counted) return fail(PC_ESTATE, "pc_export_by_owner before pc_count");
    if (c->count_kind != 0) return fail(PC_ESTATE, "pc_export_by_owner needs a global count table (pc_tune part_mode 0 or 1)");
    if (world < 1 || world > 64 || !keys_dev || !counts_dev || !offsets) return fail(PC_EINVAL, "bad export arguments");
    if (cap < (int64_t)c->h_stats.distinct) return fail(PC_EINVAL, "export buffer too small: need %llu", c->h_stats.distinct);
    StageTimer t(c, ST_EXPORT);
    CU(cudaMemsetAsync(c->d_counter.p, 0, 64 * sizeof(unsigned long long), c->stream));
    unsigned int grid = (unsigned int)std::min<uint64_t>((c->capacity + 255) / 256, 148ull * 16);
    LAUNCH(c, pc::k_owner_hist, grid, 256, c->table(), c->capacity, (unsigned int)world, c->d_counter.p);
    unsigned long long hist[64];
    READ_SMALL(c, hist, c->d_counter.p, world * sizeof(unsigned long long));
    SYNC(c);
    unsigned long long start[64]; offsets[0] = 0;
    for (int r = 0; r < world; ++r) { start[r] = (unsigned long long)offsets[r]; offsets[r + 1] = offsets[r] + (int64_t)hist[r]; }
    CU(cudaMemcpyAsync(c->d_counter.p, start, world * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, pc::k_owner_scatter, grid, 256, c->table(), c->capacity, (unsigned int)world, c->d_counter.p, keys_dev, counts_dev);
    SYNC(c);
    return PC_OK;
}

int pc_merge_counts(pc_ctx* c, const uint64_t* keys_dev, const uint32_t* counts_dev, int64_t n, int k, int64_t capacity,
                    int fresh) {
    int rc = use_device(c); if (rc) return rc;
    if (n < 0 || (n > 0 && (!keys_dev || !counts_dev))) return fail(PC_EINVAL, "bad merge arguments");
    if (fresh) {
        if (capacity <= 0) capacity = (int64_t)(n / 0.6) + 1024;
        rc = table_prepare(c, k, capacity); if (rc) return rc;
        c->pb = 0; c->part_S = (uint64_t)c->capacity;
        c->selected = c->built = c->tfidf_done = false;
    } else if (!c->counted || c->k != k) {
        return fail(PC_ESTATE, "pc_merge_counts(fresh=0) needs an existing table with the same k");
    } else if (c->pb != 0 || c->count_kind != 0) {
        return fail(PC_ESTATE, "pc_merge_counts(fresh=0) cannot add to a partitioned count table");
    }
    {
        StageTimer t(c, ST_MERGE);
        if (n > 0) LAUNCH(c, pc::k_merge, blocks_for(n, 256), 256, keys_dev, counts_dev, n, c->table(), c->capacity, c->d_stats.p);
    }
    rc = fetch_stats(c); if (rc) return rc;
    c->counted = true; c->count_kind = 0;
    return PC_OK;
}

// ---------------------------------------------------------------------------------------------- N1 subgenome extraction
int pc_sub_begin(pc_ctx* c, int n_groups) {
    int rc = use_device(c); if (rc) return rc;
    if (n_groups < 1 || n_groups > pc::kMaxGroups)
        return fail(PC_EINVAL, "n_groups = %d unsupported: 1..%d subgenomes per session", n_groups, pc::kMaxGroups);
    SYNC(c);
    c->sub_clear();
    c->sub_groups = n_groups;
    return PC_OK;
}

int pc_sub_add_counts(pc_ctx* c, int group, uint32_t min_count, int64_t* n_kmers) {
    int rc = use_device(c); if (rc) return rc;
    if (c->sub_groups == 0) return fail(PC_ESTATE, "pc_sub_add_counts before pc_sub_begin");
    if (!c->counted) return fail(PC_ESTATE, "pc_sub_add_counts before pc_count");
    if (group < 0 || group >= c->sub_groups) return fail(PC_EINVAL, "group %d out of range (have %d)", group, c->sub_groups);
    bool any = false;
    for (int g = 0; g < c->sub_groups; ++g) any |= c->sub_dict[g].set;
    if (any && c->sub_k != c->k)
        return fail(PC_ESTATE, "dictionaries of k = %d are pending: call pc_sub_compare before counting k = %d", c->sub_k, c->k);
    c->sub_k = c->k;
    if (min_count == 0) min_count = 1;
    pc_ctx::SubDict& d = c->sub_dict[group];
    StageTimer t(c, ST_SUB_DICT);
    int64_t n = 0;
    rc = select_into(c, min_count, 0xFFFFFFFFu, nullptr, nullptr, 0, &n); if (rc) return rc;
    rc = d.keys.alloc(std::max<int64_t>(n, 1)); if (rc) return rc;
    rc = d.counts.alloc(std::max<int64_t>(n, 1)); if (rc) return rc;
    int64_t n2 = 0;
    rc = select_into(c, min_count, 0xFFFFFFFFu, d.keys.p, d.counts.p, n, &n2); if (rc) return rc;
    if (n2 != n) return fail(PC_ESTATE, "count table changed while it was being read");
    d.cap = std::max<uint64_t>(64, (uint64_t)n * 4);
    if (d.cap >= (1ull << 32)) return fail(PC_EINVAL, "dictionary of %lld k-mers is too large", (long long)n);
    rc = d.table.alloc(d.cap); if (rc) return rc;
    CU(cudaMemsetAsync(d.table.p, 0, d.cap * sizeof(pc::RepSlot), c->stream));
    if (n > 0) LAUNCH(c, pc::k_kv_build, blocks_for(n, 256), 256, d.keys.p, d.counts.p, n, d.table.p, d.cap);
    d.n = n; d.set = true;
    if (n_kmers) *n_kmers = n;
    return PC_OK;
}

int pc_sub_compare(pc_ctx* c, double ratio_threshold, double default_value, int64_t sampling, int64_t* n_diff) {
    int rc = use_device(c); if (rc) return rc;
    const int G = c->sub_groups;
    if (G == 0) return fail(PC_ESTATE, "pc_sub_compare before pc_sub_begin");
    for (int g = 0; g < G; ++g)
        if (!c->sub_dict[g].set) return fail(PC_ESTATE, "subgenome %d has no dictionary for k = %d", g, c->sub_k);
    if (default_value == 0.0) return fail(PC_EINVAL, "default_kmercount_value = 0 divides by zero (the reference raises too)");
    if (c->sub_n_len >= pc::kMaxSubK) return fail(PC_EINVAL, "more than %d k-mer lengths in one session", pc::kMaxSubK);
    for (int q = 0; q < c->sub_n_len; ++q)
        if (c->sub_len[q].k == c->sub_k) return fail(PC_EINVAL, "k = %d was already compared in this session", c->sub_k);
    if (sampling < 1) sampling = 1;                            // polycracker.py:738-739
    pc_ctx::SubLen& L = c->sub_len[c->sub_n_len];
    L.k = c->sub_k;
    StageTimer t(c, ST_SUB_COMPARE);
    pc::GroupTables all{}; all.n = G;
    for (int g = 0; g < G; ++g) { all.t[g] = c->sub_dict[g].table.p; all.cap[g] = c->sub_dict[g].cap; }
    int64_t total = 0;
    for (int g = 0; g < G; ++g) {
        pc_ctx::SubDict& d = c->sub_dict[g];
        pc::GroupTables others{}; others.n = 0;
        for (int o = 0; o < G; ++o) if (o != g) { others.t[others.n] = all.t[o]; others.cap[others.n] = all.cap[o]; ++others.n; }
        rc = c->sub_tmp_a.alloc(std::max<int64_t>(d.n, 1)); if (rc) return rc;
        rc = c->sub_tmp_b.alloc(std::max<int64_t>(d.n, 1)); if (rc) return rc;
        unsigned long long nd = 0;
        if (d.n > 0) {
            CU(cudaMemsetAsync(c->d_counter.p, 0, sizeof(unsigned long long), c->stream));
            LAUNCH(c, pc::k_diff_test, blocks_for(d.n, 256), 256, d.keys.p, d.counts.p, d.n, others, ratio_threshold,
                   default_value, c->sub_tmp_a.p, c->d_counter.p);
            READ_SMALL(c, &nd, c->d_counter.p, sizeof nd);
            SYNC(c);
        }
        int64_t n = (int64_t)nd;
        uint64_t* sorted = nullptr;
        rc = radix_sort_u64(c, c->sub_tmp_a.p, c->sub_tmp_b.p, n, 2 * L.k, &sorted); if (rc) return rc;
        int64_t n_out = (n + sampling - 1) / sampling;         // every sampling-th survivor, in k-mer order
        rc = L.diff_keys[g].alloc(std::max<int64_t>(n_out, 1)); if (rc) return rc;
        rc = L.diff_counts[g].alloc(std::max<int64_t>(n_out * G, 1)); if (rc) return rc;
        if (n_out > 0) {
            LAUNCH(c, pc::k_stride_gather, blocks_for(n_out, 256), 256, sorted, n_out, sampling, L.diff_keys[g].p);
            LAUNCH(c, pc::k_diff_counts, blocks_for(n_out, 256), 256, L.diff_keys[g].p, n_out, all, L.diff_counts[g].p);
        }
        L.n_diff[g] = n_out; total += n_out;
        if (n_diff) n_diff[g] = n_out;
    }
    L.cap = std::max<uint64_t>(64, (uint64_t)total * 4);
    if (L.cap >= (1ull << 32)) return fail(PC_EINVAL, "%lld differential k-mers are too many for the mask table", (long long)total);
    rc = L.mask.alloc(L.cap); if (rc) return rc;
    CU(cudaMemsetAsync(L.mask.p, 0, L.cap * sizeof(pc::RepSlot), c->stream));
    for (int g = 0; g < G; ++g)
        if (L.n_diff[g] > 0)
            LAUNCH(c, pc::k_mask_build, blocks_for(L.n_diff[g], 256), 256, L.diff_keys[g].p, L.n_diff[g], 1 << g, L.mask.p, L.cap);
    SYNC(c);
    for (int g = 0; g < G; ++g) {                              // the dictionaries of this k are done
        pc_ctx::SubDict& d = c->sub_dict[g];
        d.table.release(); d.keys.release(); d.counts.release(); d.cap = 0; d.n = 0; d.set = false;
    }
    ++c->sub_n_len;
    return PC_OK;
}

// N4: union of the pending dictionaries, filtered like bio_hyp_class (utilities.py:255-258)
int pc_sub_union(pc_ctx* c, int64_t ratio_threshold, int64_t default_value, int64_t max_val, int64_t* n_union, int64_t* n_kept) {
    int rc = use_device(c); if (rc) return rc;
    const int G = c->sub_groups;
    if (G == 0) return fail(PC_ESTATE, "pc_sub_union before pc_sub_begin");
    for (int g = 0; g < G; ++g)
        if (!c->sub_dict[g].set) return fail(PC_ESTATE, "genome %d has no dictionary for k = %d", g, c->sub_k);
    StageTimer t(c, ST_SUB_COMPARE);
    pc::GroupTables all{}; all.n = G;
    int64_t total = 0;
    for (int g = 0; g < G; ++g) { all.t[g] = c->sub_dict[g].table.p; all.cap[g] = c->sub_dict[g].cap; total += c->sub_dict[g].n; }
    rc = c->sub_tmp_a.alloc(std::max<int64_t>(total, 1)); if (rc) return rc;
    rc = c->sub_tmp_b.alloc(std::max<int64_t>(total, 1)); if (rc) return rc;
    CU(cudaMemsetAsync(c->d_counter.p, 0, 2 * sizeof(unsigned long long), c->stream));
    for (int g = 0; g < G; ++g) {
        pc_ctx::SubDict& d = c->sub_dict[g];
        if (d.n > 0)
            LAUNCH(c, pc::k_union_test, blocks_for(d.n, 256), 256, d.keys.p, d.counts.p, d.n, g, all, (long long)ratio_threshold,
                   (long long)default_value, (long long)max_val, c->sub_tmp_a.p, c->d_counter.p, c->d_counter.p + 1);
    }
    unsigned long long h[2] = {0, 0};
    READ_SMALL(c, h, c->d_counter.p, sizeof h);
    SYNC(c);
    const int64_t n = (int64_t)h[0];
    uint64_t* sorted = nullptr;
    rc = radix_sort_u64(c, c->sub_tmp_a.p, c->sub_tmp_b.p, n, 2 * c->sub_k, &sorted); if (rc) return rc;
    rc = c->union_keys.alloc(std::max<int64_t>(n, 1)); if (rc) return rc;
    rc = c->union_counts.alloc(std::max<int64_t>(n * G, 1)); if (rc) return rc;
    if (n > 0) {
        CU(cudaMemcpyAsync(c->union_keys.p, sorted, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c->stream));
        LAUNCH(c, pc::k_diff_counts, blocks_for(n, 256), 256, c->union_keys.p, n, all, c->union_counts.p);
    }
    SYNC(c);
    c->n_union_kept = n;
    if (n_union) *n_union = (int64_t)h[1];
    if (n_kept) *n_kept = n;
    return PC_OK;
}

int pc_sub_union_get(pc_ctx* c, uint64_t* keys, uint32_t* counts, int loc, int64_t* n) {
    int rc = use_device(c); if (rc) return rc;
    if (c->sub_groups == 0 || c->n_union_kept < 0) return fail(PC_ESTATE, "pc_sub_union_get before pc_sub_union");
    if (n) *n = c->n_union_kept;
    if (keys) { rc = copy_out(c, keys, c->union_keys.p, (size_t)c->n_union_kept * sizeof(uint64_t), loc); if (rc) return rc; }
    if (counts) { rc = copy_out(c, counts, c->union_counts.p, (size_t)c->n_union_kept * c->sub_groups * sizeof(uint32_t), loc); if (rc) return rc; }
    return PC_OK;
}

int pc_sub_diff_get(pc_ctx* c, int k, int group, uint64_t* keys, uint32_t* counts, int loc, int64_t* n) {
    int rc = use_device(c); if (rc) return rc;
    if (group < 0 || group >= c->sub_groups) return fail(PC_EINVAL, "group %d out of range (have %d)", group, c->sub_groups);
    for (int q = 0; q < c->sub_n_len; ++q) {
        pc_ctx::SubLen& L = c->sub_len[q];
        if (L.k != k) continue;
        if (n) *n = L.n_diff[group];
        if (keys) { rc = copy_out(c, keys, L.diff_keys[group].p, L.n_diff[group] * sizeof(uint64_t), loc); if (rc) return rc; }
        if (counts) { rc = copy_out(c, counts, L.diff_counts[group].p, L.n_diff[group] * c->sub_groups * sizeof(uint32_t), loc); if (rc) return rc; }
        return PC_OK;
    }
    return fail(PC_ESTATE, "k = %d has not been compared in this session", k);
}

int pc_sub_mapback(pc_ctx* c, int32_t* coverage, int loc) {
    int rc = use_device(c); if (rc) return rc;
    if (c->sub_groups == 0 || c->sub_n_len == 0) return fail(PC_ESTATE, "pc_sub_mapback before pc_sub_compare");
    if (!c->loaded) return fail(PC_ESTATE, "pc_sub_mapback before pc_load_sequences");
    if (!coverage) return fail(PC_EINVAL, "coverage is null");
    const int G = c->sub_groups;
    rc = c->d_cov.alloc(std::max<int64_t>(c->n_chunks * G, 1)); if (rc) return rc;
    {
        StageTimer t(c, ST_SUB_MAPBACK);
        CU(cudaMemsetAsync(c->d_cov.p, 0, std::max<int64_t>(c->n_chunks * G, 1) * sizeof(int), c->stream));
        pc::SubKSet ks{}; ks.n = c->sub_n_len;
        for (int q = 0; q < ks.n; ++q) { ks.t[q] = c->sub_len[q].mask.p; ks.cap[q] = c->sub_len[q].cap; ks.k[q] = c->sub_len[q].k; }
        if (c->n_words > 0)
            LAUNCH(c, pc::k_probe_groups<4>, blocks_for(c->n_words, 256), 256, c->words.p, c->nmask.p, c->n_words, ks,
                   c->d_word0.p, c->n_chunks, G, c->d_cov.p);
    }
    return copy_out(c, coverage, c->d_cov.p, (size_t)c->n_chunks * G * sizeof(int32_t), loc);
}

// ---------------------------------------------------------------------------------------------- K4
int pc_select(pc_ctx* c, uint32_t low, uint32_t high, int use_high, int64_t sampling, int64_t* n_selected) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->counted) return fail(PC_ESTATE, "pc_select before pc_count");
    if (sampling < 1) sampling = 1;                            // polycracker.py:220-221
    if (!use_high) high = 0xFFFFFFFFu;
    if (low == 0) low = 1;                                     // absent k-mers are never selected
    c->selected = c->built = c->tfidf_done = false;
    int64_t cap = std::min<int64_t>((int64_t)c->h_stats.distinct, (int64_t)1 << 26);
    cap = std::max<int64_t>(cap, 1024);
    int64_t n = 0;
    if (c->cross_valid && !use_high && c->cross_low == low) {
        // K4 was fused into the count: the k-mers that reached `low` are already listed
        StageTimer t(c, ST_SELECT);
        unsigned long long nc = 0;
        READ_SMALL(c, &nc, c->d_cross_n.p, sizeof nc);
        SYNC(c);
        n = (int64_t)nc;
        rc = c->sel_a.alloc(std::max<int64_t>(n, 1024)); if (rc) return rc;
        if (n) CU(cudaMemcpyAsync(c->sel_a.p, c->cross_keys.p, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c->stream));
    } else {
        TRACE("select: begin");
        StageTimer t(c, ST_SELECT);
        rc = c->sel_a.alloc(cap); if (rc) return rc;
        TRACE("select: alloc done");
        rc = select_into(c, low, high, c->sel_a.p, nullptr, cap, &n); if (rc) return rc;
        TRACE("select: scan done");
        if (n > cap) {                                         // rare: more survivors than the first guess
            cap = n;
            rc = c->sel_a.alloc(cap); if (rc) return rc;
            rc = select_into(c, low, high, c->sel_a.p, nullptr, cap, &n); if (rc) return rc;
        }
    }
    if (c->select_raw && sampling <= 1) {
        // multi-GPU owners (pc_tune select_raw 1): the survivors leave unsorted -- every rank sorts the all-gathered set in
        // pc_set_selected anyway, so the owner's own sort and repeat table (0.45 ms per step) would be thrown away
        c->sel = c->sel_a.p; c->n_sel = n;
        c->selected = true; c->sel_unsorted = true;
        if (n_selected) *n_selected = n;
        return PC_OK;
    }
    c->sel_unsorted = false;
    {
        StageTimer t(c, ST_SORT);
        rc = c->sel_b.alloc(std::max<int64_t>(n, 1)); if (rc) return rc;
        uint64_t* sorted = nullptr;
        rc = radix_sort_u64(c, c->sel_a.p, c->sel_b.p, n, 2 * c->k, &sorted); if (rc) return rc;
        if (sampling > 1 && n > 0) {
            int64_t n_out = (n + sampling - 1) / sampling;     // survivors 0, s, 2s, ... (count % s == 0)
            uint64_t* other = (sorted == c->sel_a.p) ? c->sel_b.p : c->sel_a.p;
            LAUNCH(c, pc::k_stride_gather, blocks_for(n_out, 256), 256, sorted, n_out, sampling, other);
            sorted = other; n = n_out;
        }
        c->sel = sorted; c->n_sel = n;
    }
    TRACE("select: sort queued");
    rc = build_rep_table(c); if (rc) return rc;
    TRACE("select: rep table queued");
    c->selected = true;
    if (n_selected) *n_selected = n;
    return PC_OK;
}

int pc_selected_get(pc_ctx* c, uint64_t* keys, int loc) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->selected) return fail(PC_ESTATE, "no repeat set selected");
    return copy_out(c, keys, c->sel, c->n_sel * sizeof(uint64_t), loc);
}

int pc_set_selected(pc_ctx* c, const uint64_t* keys, int64_t n, int k, int loc, int64_t* n_unique) {
    int rc = use_device(c); if (rc) return rc;
    if (k < 1 || k > 32) return fail(PC_EINVAL, "k = %d unsupported: this path handles 1 <= k <= 32 (uint64 k-mers)", k);
    if (n < 0 || (n > 0 && !keys)) return fail(PC_EINVAL, "bad set_selected arguments");
    if (c->counted && c->k != k) return fail(PC_EINVAL, "k = %d differs from the count table's k = %d", k, c->k);
    c->k = k;
    c->selected = c->built = c->tfidf_done = false;
    StageTimer t(c, ST_SORT);
    rc = c->sel_a.alloc(std::max<int64_t>(n, 1)); if (rc) return rc;
    rc = c->sel_b.alloc(std::max<int64_t>(n, 1)); if (rc) return rc;
    if (n) CU(cudaMemcpyAsync(c->sel_a.p, keys, n * sizeof(uint64_t), loc == PC_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    if (n) LAUNCH(c, pc::k_canonicalize, blocks_for(n, 256), 256, c->sel_a.p, n, k);
    uint64_t* sorted = nullptr;
    rc = radix_sort_u64(c, c->sel_a.p, c->sel_b.p, n, 2 * k, &sorted); if (rc) return rc;
    uint64_t* other = (sorted == c->sel_a.p) ? c->sel_b.p : c->sel_a.p;
    unsigned long long nu = 0;
    uint64_t* result = sorted;
    if (n > 0) {
        // entries that are not k-mers (the sentinel pc_kmer_encode returns for a name with N or any non-ACGT character)
        // are dropped: canonicalising them would turn every one into poly-A and match windows the query never named
        unsigned long long sent = 0;
        CU(cudaMemsetAsync(c->d_counter.p, 0, 2 * sizeof(unsigned long long), c->stream));
        LAUNCH(c, pc::k_count_sentinel, blocks_for(n, 256), 256, sorted, n, c->d_counter.p + 1);
        READ_SMALL(c, &sent, c->d_counter.p + 1, sizeof sent);
        SYNC(c);
        n -= (int64_t)sent;                                    // sorted last
    }
    if (n > 0) {
        unsigned long long dups = 0;
        LAUNCH(c, pc::k_count_dups, blocks_for(n, 256), 256, sorted, n, c->d_counter.p);
        READ_SMALL(c, &dups, c->d_counter.p, sizeof dups);
        SYNC(c);
        nu = (unsigned long long)n;
        if (dups) {                                            // rare: a k-mer FASTA listing a k-mer (or both strands) twice
            LAUNCH(c, pc::k_unique_sorted_single, 1, 1024, sorted, n, other, c->d_counter.p);
            READ_SMALL(c, &nu, c->d_counter.p, sizeof nu);
            SYNC(c);
            result = other;
        }
    }
    c->sel = result; c->n_sel = (int64_t)nu; c->sel_unsorted = false;
    rc = build_rep_table(c); if (rc) return rc;
    c->selected = true;
    if (n_unique) *n_unique = c->n_sel;
    return PC_OK;
}

// ---------------------------------------------------------------------------------------------- K5 + K6
// K6: sort every row's hit list, run-length encode it into CSR and accumulate df.  Expects hits_a / d_row_base /
// d_row_cursor / d_df / d_indptr prepared on the device and c->n_sel = number of columns.
// bitmap plan: columns per range (multiple of 1024) and the shared count slots that fit one CTA's 227 KB next to the
// bitmap (4 B per 32 columns), its per-word prefixes (2 B per 32 columns) and per-chunk prefixes
struct CsrPlan { bool bitmap; unsigned int range_bits, range_shift, ncap, cnt_alloc, n_ranges; size_t smem; };
static CsrPlan csr_plan(const pc_ctx* c) {
    CsrPlan p{};
    const int64_t budget = 231424;                             // dynamic shared memory (static: < 2 KB)
    const int64_t n_cols = std::max<int64_t>(c->n_sel, 1);
    int64_t rb = ((n_cols + 1023) / 1024) * 1024;
    if (c->csr_range_bits > 0) rb = std::min<int64_t>(rb, c->csr_range_bits);
    if (rb > 768 * 1024 || rb < n_cols) {
        // several ranges: a power of two (the range of a column is a shift), at most 512 K columns = 64 KB of bits + 32 KB
        // of prefixes, which leaves 32 k count slots
        int64_t pow2 = 1024;
        while (pow2 * 2 <= std::min<int64_t>(rb, 512 * 1024)) pow2 *= 2;
        rb = pow2;
    }
    unsigned int shift = 0; while (((int64_t)1 << shift) < rb) ++shift;
    const int64_t nchunks = rb / 1024, NC = nchunks | 1;
    const int64_t fixed = 32 * NC * 4 + (nchunks + 1) * 4 + 32 * NC * 2;
    const int64_t n_ranges = (n_cols + rb - 1) / rb;
    const int64_t cnt_alloc = (budget - fixed) / 4;
    int64_t ncap = cnt_alloc;
    if (c->csr_ncap > 0) ncap = std::min<int64_t>(ncap, c->csr_ncap);
    p.range_bits = (unsigned int)rb; p.range_shift = shift; p.ncap = (unsigned int)ncap; p.cnt_alloc = (unsigned int)cnt_alloc;
    p.n_ranges = (unsigned int)n_ranges;
    p.smem = (size_t)(fixed + cnt_alloc * 4);
    const bool fits = n_ranges <= PC_CSR_MAXR && 32 * n_ranges <= cnt_alloc;
    // auto: the bitmap while the repeat set spans at most 6 ranges (3 M columns), the per-row radix sort beyond: every
    // range of every row pays its own prefix passes and barriers (measured at cfg2, bitmap / sort: 0.67 M columns
    // 1.17 / 2.39 ms, 2.5 M (5 ranges) 3.32 / 3.97, 6.8 M (14 ranges) 7.05 / 4.78)
    p.bitmap = fits && (c->csr_mode == 1 || (c->csr_mode < 0 && n_ranges <= 6));
    return p;
}

static int rows_to_csr_bitmap(pc_ctx* c, int64_t n_rows, const CsrPlan& plan) {
    int rc;
    c->nnz = 0;
    if (n_rows > 0) {
        StageTimer t(c, ST_ROWSORT);
        // capacity of indices / data: a row has at most as many entries as hits.  The exact hit total costs one tiny kernel
        // and a read-back (~30 us); sizing by the rows' windows instead needs no read-back but 36 GB at 4.5 Gbp, which no
        // longer fits next to the k-mer buffers of a following k < 17 pass on one GPU.
        unsigned long long total_hits = 0;
        CU(cudaMemsetAsync(c->d_counter.p, 0, 2 * sizeof(unsigned long long), c->stream));
        LAUNCH(c, pc::k_sum_u32, (unsigned int)std::min<int64_t>(148 * 4, (n_rows + 255) / 256), 256, c->d_row_cursor.p, n_rows, c->d_counter.p);
        READ_SMALL(c, &total_hits, c->d_counter.p, sizeof total_hits);
        SYNC(c);
        rc = c->d_indices.alloc(std::max<size_t>((size_t)total_hits, 1)); if (rc) return rc;
        rc = c->d_data.alloc(std::max<size_t>((size_t)total_hits, 1)); if (rc) return rc;
        if (plan.n_ranges > 1) { rc = c->hits_b.alloc(std::max<size_t>(c->hits_a.n, 1)); if (rc) return rc; }
        rc = c->d_row_state.alloc(n_rows + 1); if (rc) return rc;
        CU(cudaMemsetAsync(c->d_row_state.p, 0, (n_rows + 1) * sizeof(unsigned long long), c->stream));
        unsigned int* queue = reinterpret_cast<unsigned int*>(c->d_row_state.p + n_rows);
        CU(cudaFuncSetAttribute(pc::k_row_bitmap_csr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
        c->launches++;
        pc::k_row_bitmap_csr<<<(unsigned int)std::min<int64_t>(148, n_rows), PC_CSR_THREADS, plan.smem, c->stream>>>(
            c->hits_a.p, c->d_row_base.p, c->d_row_cursor.p, (int)n_rows, (unsigned int)c->n_sel, plan.range_bits, plan.ncap,
            c->d_row_state.p, queue, c->d_indptr.p, c->d_indices.p, c->d_data.p, c->d_df.p,
            plan.n_ranges > 1 ? c->hits_b.p : nullptr, plan.range_shift, plan.cnt_alloc);
        CU(cudaGetLastError());
        int64_t total = 0;
        READ_SMALL(c, &total, c->d_indptr.p + n_rows, sizeof(int64_t));
        SYNC(c);
        c->nnz = total;
    }
    SYNC(c);                      // host staging vectors no longer needed
    c->n_rows = n_rows;
    c->built = true;
    return PC_OK;
}

static int rows_to_csr(pc_ctx* c, int64_t n_rows) {
    int rc;
    {
        const CsrPlan plan = csr_plan(c);
        if (plan.bitmap && c->n_sel > 0) return rows_to_csr_bitmap(c, n_rows, plan);
    }
    rc = c->hits_b.alloc(std::max<int64_t>(c->hits_a.n, 1)); if (rc) return rc;
    rc = c->d_row_nnz.alloc(std::max<int64_t>(n_rows, 1)); if (rc) return rc;
    int bits = 1; while (((int64_t)1 << bits) < c->n_sel) ++bits;
    // digit width: fewest passes, then the narrowest digit (smem = 17 * 2^DB * 4 bytes)
    int db = 8;
    {
        int best = (bits + 7) / 8;
        if ((bits + 9) / 10 < best) { db = 10; best = (bits + 9) / 10; }
        if ((bits + 10) / 11 < best) { db = 11; best = (bits + 10) / 11; }
    }
    int n_passes = std::max(1, (bits + db - 1) / db);
    int32_t* sorted = (n_passes & 1) ? c->hits_b.p : c->hits_a.p;
    c->nnz = 0;
    if (n_rows > 0) {
        {
            StageTimer t(c, ST_ROWSORT);
            c->launches++;
#define ROW_SORT(W_, DB_)                                                                                                  \
            do {                                                                                                           \
                size_t smem = (size_t)((W_) + 1) * ((size_t)1 << (DB_)) * sizeof(unsigned int);                            \
                CU(cudaFuncSetAttribute((pc::k_row_sort<W_, DB_>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                pc::k_row_sort<W_, DB_><<<(unsigned int)n_rows, (W_) * 32, smem, c->stream>>>(                            \
                    c->hits_a.p, c->hits_b.p, c->d_row_base.p, c->d_row_cursor.p, n_passes, c->d_row_nnz.p);               \
            } while (0)
            // per-warp histograms: 16 warps x 2^DB counters are zeroed, scanned and offset every pass, which is as much
            // shared-memory work as ranking a 25 k-hit row; rowsort_warps (8 or 16) trades that against parallelism
            if (c->rowsort_warps == 8) {
                if (db == 8) ROW_SORT(8, 8); else if (db == 10) ROW_SORT(8, 10); else ROW_SORT(8, 11);
            } else {
                if (db == 8) ROW_SORT(16, 8); else if (db == 10) ROW_SORT(16, 10); else ROW_SORT(16, 11);
            }
#undef ROW_SORT
            CU(cudaGetLastError());
            LAUNCH(c, pc::k_indptr, 1, 1024, c->d_row_nnz.p, n_rows, c->d_indptr.p);
        }
        int64_t total = 0;
        READ_SMALL(c, &total, c->d_indptr.p + n_rows, sizeof(int64_t));
        SYNC(c);
        c->nnz = total;
        rc = c->d_indices.alloc(std::max<int64_t>(total, 1)); if (rc) return rc;
        rc = c->d_data.alloc(std::max<int64_t>(total, 1)); if (rc) return rc;
        StageTimer t(c, ST_EMIT);
        LAUNCH(c, (pc::k_row_emit<256>), (unsigned int)n_rows, 256, sorted, c->d_row_base.p, c->d_row_cursor.p,
               c->d_indptr.p, c->d_indices.p, c->d_data.p, c->d_df.p);
    }
    SYNC(c);                      // host staging vectors no longer needed
    c->n_rows = n_rows;
    c->built = true;
    return PC_OK;
}

int pc_build_matrix(pc_ctx* c, const int32_t* chunk_row, int64_t n_rows, int64_t* nnz) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->loaded || !c->selected) return fail(PC_ESTATE, "pc_build_matrix needs loaded sequences and a repeat set");
    if (c->sel_unsorted) return fail(PC_ESTATE, "the repeat set was selected raw (pc_tune select_raw): install the gathered set with pc_set_selected first");
    if (n_rows < 0 || (c->n_chunks > 0 && !chunk_row)) return fail(PC_EINVAL, "bad build_matrix arguments");
    if (n_rows > 0x7FFFFFFF) return fail(PC_EINVAL, "too many rows");
    c->built = c->tfidf_done = false; c->n_esc = -1; c->n_esc8_col = c->n_esc8_cnt = -1;
    const int k = c->k;
    std::vector<int64_t> row_base(n_rows + 1, 0), row_windows(std::max<int64_t>(n_rows, 1), -1);
    for (int64_t i = 0; i < c->n_chunks; ++i) {
        int32_t r = chunk_row[i];
        if (r < 0) continue;
        if (r >= n_rows) return fail(PC_EINVAL, "chunk_row[%lld] = %d >= n_rows", (long long)i, r);
        if (row_windows[r] >= 0) return fail(PC_EINVAL, "row %d is assigned to two chunks", r);
        row_windows[r] = std::max<int64_t>(0, c->h_len[i] - k + 1);
    }
    for (int64_t r = 0; r < n_rows; ++r) {
        int64_t wnd = std::max<int64_t>(0, row_windows[r]);
        if (wnd > 0xFFFFFFFFll) return fail(PC_EINVAL, "chunk too long for 32-bit hit cursors");
        row_base[r + 1] = row_base[r] + wnd;
    }
    int64_t total_windows = row_base[n_rows];
    rc = c->d_chunk_row.alloc(c->n_chunks); if (rc) return rc;
    rc = c->d_row_base.alloc(n_rows + 1); if (rc) return rc;
    rc = c->d_row_cursor.alloc(n_rows); if (rc) return rc;
    rc = c->d_indptr.alloc(n_rows + 1); if (rc) return rc;
    rc = c->hits_a.alloc(total_windows); if (rc) return rc;
    rc = c->d_df.alloc(std::max<int64_t>(c->n_sel, 1)); if (rc) return rc;
    if (c->n_chunks) CU(cudaMemcpyAsync(c->d_chunk_row.p, chunk_row, c->n_chunks * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_row_base.p, row_base.data(), (n_rows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(c->d_row_cursor.p, 0, std::max<int64_t>(n_rows, 1) * sizeof(unsigned int), c->stream));
    CU(cudaMemsetAsync(c->d_df.p, 0, std::max<int64_t>(c->n_sel, 1) * sizeof(int), c->stream));
    CU(cudaMemsetAsync(c->d_indptr.p, 0, (n_rows + 1) * sizeof(int64_t), c->stream));
    {
        StageTimer t(c, ST_PROBE);
        if (c->n_words > 0 && n_rows > 0 && c->n_sel > 0) {
            bool use_stream = c->stream_valid && c->stream_k == k && c->probe_mode != 0;
            if (c->probe_mode == 1 && !use_stream) return fail(PC_ESTATE, "probe_mode=1 needs the k-mer stream of a partitioned pc_count");
            unsigned int grid = blocks_for(c->n_words, 256);
            // fingerprint-first walk when the 16 B/slot table is far beyond L2 (measured: +17 % at 43 MB, -9 % at 650 MB)
            const bool use_fp = c->probe_fp == 1 || (c->probe_fp < 0 && use_stream && c->rcap * sizeof(pc::RepSlot) > (size_t)(256ull << 20));
            if (use_stream && c->rfilter_bits > 0 && !use_fp) {
                // one persistent CTA of 1024 threads per SM, or two of 512 when two copies of the bitmap fit
                size_t smem = (size_t)(c->rfilter_bits / 32 + 4) * sizeof(unsigned int);
                const bool two = c->probe_threads == 512 && smem <= (size_t)(110 << 10);   // measured slower than one CTA of 1024
                const int threads = two ? 512 : 1024;
                unsigned int pgrid = (unsigned int)std::min<int64_t>(two ? 296 : 148, (c->n_words + threads - 1) / threads);
                c->launches++;
                if (two) {
                    CU(cudaFuncSetAttribute(pc::k_probe_stream_filt<4, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    pc::k_probe_stream_filt<4, 512><<<pgrid, 512, smem, c->stream>>>(c->kstream.p, c->n_words, c->d_word0.p, c->n_chunks,
                        c->d_chunk_row.p, c->d_row_base.p, c->rtable.p, c->rcap, c->rfilter.p, c->rfilter_bits, c->hits_a.p, c->d_row_cursor.p);
                } else {
                    CU(cudaFuncSetAttribute(pc::k_probe_stream_filt<4, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    pc::k_probe_stream_filt<4, 1024><<<pgrid, 1024, smem, c->stream>>>(c->kstream.p, c->n_words, c->d_word0.p, c->n_chunks,
                        c->d_chunk_row.p, c->d_row_base.p, c->rtable.p, c->rcap, c->rfilter.p, c->rfilter_bits, c->hits_a.p, c->d_row_cursor.p);
                }
                CU(cudaGetLastError());
            } else if (use_stream) {
                if (c->probe_batch == 4)
                    LAUNCH(c, (pc::k_probe_stream<4>), grid, 256, c->kstream.p, c->n_words, c->d_word0.p, c->n_chunks,
                           c->d_chunk_row.p, c->d_row_base.p, c->rtable.p, c->rcap, use_fp ? c->rfp.p : nullptr,
                           c->gfilter_bits ? c->rfilter.p : nullptr, c->gfilter_bits, c->hits_a.p, c->d_row_cursor.p);
                else
                    LAUNCH(c, (pc::k_probe_stream<8>), grid, 256, c->kstream.p, c->n_words, c->d_word0.p, c->n_chunks,
                           c->d_chunk_row.p, c->d_row_base.p, c->rtable.p, c->rcap, use_fp ? c->rfp.p : nullptr,
                           c->gfilter_bits ? c->rfilter.p : nullptr, c->gfilter_bits, c->hits_a.p, c->d_row_cursor.p);
            } else if (c->probe_roll == 0) {
                LAUNCH(c, (pc::k_probe<8>), grid, 256, c->words.p, c->nmask.p, c->n_words, k, c->d_word0.p, c->n_chunks,
                       c->d_chunk_row.p, c->d_row_base.p, c->rtable.p, c->rcap, use_fp ? c->rfp.p : nullptr, c->hits_a.p, c->d_row_cursor.p);
            } else if (c->rfilter_bits > 0 && !use_fp) {
                // no k-mer stream (super-k-mer count): roller + shared-memory bitmap, one persistent CTA of 1024 threads per SM
                size_t smem = (size_t)(c->rfilter_bits / 32 + 4) * sizeof(unsigned int);
                unsigned int pgrid = (unsigned int)std::min<int64_t>(148, (c->n_words + 1023) / 1024);
                c->launches++;
                if (c->rcap8 > 0) {
                    CU(cudaFuncSetAttribute((pc::k_probe_roll_filt<4, 1024, true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    pc::k_probe_roll_filt<4, 1024, true, true><<<pgrid, 1024, smem, c->stream>>>(c->words.p, c->nmask.p, c->n_words, k, c->d_word0.p,
                        c->n_chunks, c->d_chunk_row.p, c->d_row_base.p, c->rtable.p, c->rcap, nullptr, c->rfilter.p, c->rfilter_bits,
                        c->rtable8.p, c->rcap8, c->sel, c->hits_a.p, c->d_row_cursor.p);
                } else {
                    CU(cudaFuncSetAttribute((pc::k_probe_roll_filt<4, 1024, true, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    pc::k_probe_roll_filt<4, 1024, true, false><<<pgrid, 1024, smem, c->stream>>>(c->words.p, c->nmask.p, c->n_words, k, c->d_word0.p,
                        c->n_chunks, c->d_chunk_row.p, c->d_row_base.p, c->rtable.p, c->rcap, nullptr, c->rfilter.p, c->rfilter_bits,
                        nullptr, 0, nullptr, c->hits_a.p, c->d_row_cursor.p);
                }
                CU(cudaGetLastError());
            } else {
                // (64 registers: 4 CTAs of 256 threads per SM; the first version ran BATCH = 8 at 136 registers and one CTA
                // per SM: 9.1 ms instead of ~4 for 1.3 M repeat k-mers)
                unsigned int pgrid = (unsigned int)std::min<int64_t>(148 * 4, (c->n_words + 255) / 256);
                if (c->rcap8 > 0)
                    LAUNCH(c, (pc::k_probe_roll_filt<4, 256, false, true>), pgrid, 256, c->words.p, c->nmask.p, c->n_words, k, c->d_word0.p,
                           c->n_chunks, c->d_chunk_row.p, c->d_row_base.p, c->rtable.p, c->rcap, nullptr,
                           c->gfilter_bits ? c->rfilter.p : nullptr, c->gfilter_bits, c->rtable8.p, c->rcap8, c->sel,
                           c->hits_a.p, c->d_row_cursor.p);
                else
                    LAUNCH(c, (pc::k_probe_roll_filt<4, 256, false, false>), pgrid, 256, c->words.p, c->nmask.p, c->n_words, k, c->d_word0.p,
                           c->n_chunks, c->d_chunk_row.p, c->d_row_base.p, c->rtable.p, c->rcap, use_fp ? c->rfp.p : nullptr,
                           c->gfilter_bits ? c->rfilter.p : nullptr, c->gfilter_bits, nullptr, 0, nullptr, c->hits_a.p, c->d_row_cursor.p);
            }
        }
    }
    rc = rows_to_csr(c, n_rows); if (rc) return rc;
    if (nnz) *nnz = c->nnz;
    return PC_OK;
}

// K6 alone, for hit lists that already exist as text artefacts: the separately invoked generate_Kmer_Matrix stage
// reads blasted_merged.bed (one line per chunk, comma-joined k-mer names with multiplicity, polycracker.py:343-353),
// the host maps names to ids, and the Counter/dok/tocsc work (sort + run-length encode per row) runs on the GPU.
int pc_matrix_from_hits(pc_ctx* c, const int64_t* row_offsets, const int32_t* cols, int64_t n_rows, int64_t n_cols,
                        int64_t* nnz) {
    int rc = use_device(c); if (rc) return rc;
    if (n_rows < 0 || n_cols < 0 || !row_offsets || n_rows > 0x7FFFFFFF) return fail(PC_EINVAL, "bad matrix_from_hits arguments");
    int64_t total = row_offsets[n_rows];
    if (total < 0 || (total > 0 && !cols)) return fail(PC_EINVAL, "bad hit list");
    std::vector<unsigned int> cursor(std::max<int64_t>(n_rows, 1), 0);
    for (int64_t r = 0; r < n_rows; ++r) {
        int64_t n = row_offsets[r + 1] - row_offsets[r];
        if (n < 0 || n > 0xFFFFFFFFll) return fail(PC_EINVAL, "row %lld has a bad hit count", (long long)r);
        cursor[r] = (unsigned int)n;
    }
    for (int64_t i = 0; i < total; ++i)
        if (cols[i] < 0 || (int64_t)cols[i] >= n_cols) return fail(PC_EINVAL, "hit %lld names column %d, the matrix has %lld", (long long)i, cols[i], (long long)n_cols);
    c->built = c->tfidf_done = false;
    c->n_sel = n_cols;                                         // column universe of this matrix
    rc = c->d_row_base.alloc(n_rows + 1); if (rc) return rc;
    rc = c->d_row_cursor.alloc(n_rows); if (rc) return rc;
    rc = c->d_indptr.alloc(n_rows + 1); if (rc) return rc;
    rc = c->hits_a.alloc(total); if (rc) return rc;
    rc = c->d_df.alloc(std::max<int64_t>(n_cols, 1)); if (rc) return rc;
    CU(cudaMemcpyAsync(c->d_row_base.p, row_offsets, (n_rows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    if (n_rows) CU(cudaMemcpyAsync(c->d_row_cursor.p, cursor.data(), n_rows * sizeof(unsigned int), cudaMemcpyHostToDevice, c->stream));
    if (total) CU(cudaMemcpyAsync(c->hits_a.p, cols, total * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(c->d_df.p, 0, std::max<int64_t>(n_cols, 1) * sizeof(int), c->stream));
    CU(cudaMemsetAsync(c->d_indptr.p, 0, (n_rows + 1) * sizeof(int64_t), c->stream));
    rc = rows_to_csr(c, n_rows); if (rc) return rc;
    if (nnz) *nnz = c->nnz;
    return PC_OK;
}

int pc_csr_get(pc_ctx* c, int64_t* indptr, int32_t* indices, float* data, int loc) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->built) return fail(PC_ESTATE, "no matrix built");
    StageTimer t(c, ST_D2H);
    if (indptr) { rc = copy_out(c, indptr, c->d_indptr.p, (c->n_rows + 1) * sizeof(int64_t), loc); if (rc) return rc; }
    if (indices) { rc = copy_out(c, indices, c->d_indices.p, c->nnz * sizeof(int32_t), loc); if (rc) return rc; }
    if (data) { rc = copy_out(c, data, c->d_data.p, c->nnz * sizeof(float), loc); if (rc) return rc; }
    return PC_OK;
}

// compact hand-off (6 bytes per entry instead of 12 with the float32 tf-idf array): counts as uint16 + escapes
int pc_csr_get16(pc_ctx* c, int64_t* indptr, int32_t* indices, uint16_t* data16, int loc, int64_t* n_escapes) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->built) return fail(PC_ESTATE, "no matrix built");
    if (!n_escapes) return fail(PC_EINVAL, "n_escapes is null");
    StageTimer t(c, ST_D2H);
    // an entry >= 65535 needs a chunk with >= 65535 windows holding ONE k-mer: at most one per 65535 hits
    int64_t hits = 0;
    const int64_t esc_cap = c->nnz / 65535 + c->n_rows + 16;
    rc = c->d_data16.alloc(std::max<int64_t>(c->nnz, 1)); if (rc) return rc;
    rc = c->d_esc_pos.alloc(esc_cap); if (rc) return rc;
    rc = c->d_esc_val.alloc(esc_cap); if (rc) return rc;
    (void)hits;
    unsigned long long ne = 0;
    CU(cudaMemsetAsync(c->d_counter.p + 8, 0, sizeof(unsigned long long), c->stream));
    if (c->nnz > 0)
        LAUNCH(c, pc::k_data_u16, (unsigned int)std::min<int64_t>(148 * 8, (c->nnz + 255) / 256), 256, c->d_data.p, c->nnz, c->d_data16.p,
               c->d_counter.p + 8, c->d_esc_pos.p, c->d_esc_val.p, (unsigned long long)esc_cap);
    READ_SMALL(c, &ne, c->d_counter.p + 8, sizeof ne);
    if (indptr) { rc = copy_out(c, indptr, c->d_indptr.p, (c->n_rows + 1) * sizeof(int64_t), loc); if (rc) return rc; }
    if (indices) { rc = copy_out(c, indices, c->d_indices.p, c->nnz * sizeof(int32_t), loc); if (rc) return rc; }
    if (data16) { rc = copy_out(c, data16, c->d_data16.p, c->nnz * sizeof(uint16_t), loc); if (rc) return rc; }
    SYNC(c);
    if ((int64_t)ne > esc_cap) return fail(PC_ESTATE, "%llu matrix entries >= 65535 for %lld escape slots", ne, (long long)esc_cap);
    c->n_esc = (int64_t)ne;
    *n_escapes = c->n_esc;
    return PC_OK;
}

// densest hand-off: 2 bytes per entry (column gap + count, one byte each) + escape lists
static int csr_get_packed(pc_ctx* c, int64_t* indptr, void* dcol, uint8_t* cnt8, int loc, int64_t n_escapes[2], int gap_bytes);
int pc_csr_get8(pc_ctx* c, int64_t* indptr, uint8_t* dcol8, uint8_t* cnt8, int loc, int64_t n_escapes[2]) {
    return csr_get_packed(c, indptr, dcol8, cnt8, loc, n_escapes, 1);
}
int pc_csr_get_gap16(pc_ctx* c, int64_t* indptr, uint16_t* dcol16, uint8_t* cnt8, int loc, int64_t n_escapes[2]) {
    return csr_get_packed(c, indptr, dcol16, cnt8, loc, n_escapes, 2);
}
int pc_csr_get_g12c4(pc_ctx* c, int64_t* indptr, uint16_t* packed, int loc, int64_t n_escapes[2]) {
    return csr_get_packed(c, indptr, packed, nullptr, loc, n_escapes, 3);       // 3 = 12-bit gap + 4-bit count in one uint16
}
static int csr_get_packed(pc_ctx* c, int64_t* indptr, void* dcol8, uint8_t* cnt8, int loc, int64_t n_escapes[2], int gap_bytes) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->built) return fail(PC_ESTATE, "no matrix built");
    if (!n_escapes) return fail(PC_EINVAL, "n_escapes is null");
    StageTimer t(c, ST_D2H);
    // column escapes: the first entry of every row + gaps >= 255 (at most n_selected / 255 per row); count escapes: entries
    // >= 255 (a k-mer that fills >= 255 windows of one chunk)
    // (sized so that they cannot overflow: low-complexity genomes put a count >= 255 into most entries)
    const bool g12c4 = gap_bytes == 3;
    if (g12c4) gap_bytes = 2;
    const int64_t esc_gap = g12c4 ? 4095 : gap_bytes == 1 ? 255 : 65535;
    const int64_t col_cap = c->n_rows * (1 + c->n_sel / esc_gap) + 16, cnt_cap = c->nnz + 16;
    const int64_t col_cap_used = std::min<int64_t>(col_cap, c->nnz + 16);
    rc = c->d_dcol8.alloc(std::max<int64_t>(c->nnz * gap_bytes, 1)); if (rc) return rc;
    rc = c->d_cnt8.alloc(std::max<int64_t>(c->nnz, 1)); if (rc) return rc;
    rc = c->d_esc8_pos.alloc(col_cap_used); if (rc) return rc;
    rc = c->d_esc8_val.alloc(col_cap_used); if (rc) return rc;
    rc = c->d_esc_pos.alloc(cnt_cap); if (rc) return rc;
    rc = c->d_esc_val.alloc(cnt_cap); if (rc) return rc;
    unsigned long long ne[2] = {0, 0};
    CU(cudaMemsetAsync(c->d_counter.p + 8, 0, 2 * sizeof(unsigned long long), c->stream));
    if (c->nnz > 0 && c->n_rows > 0) {
        if (g12c4)
            LAUNCH(c, (pc::k_csr_pack_g12c4<256>), (unsigned int)c->n_rows, 256, c->d_indptr.p, c->d_indices.p, c->d_data.p,
                   reinterpret_cast<unsigned short*>(c->d_dcol8.p), c->d_counter.p + 8, c->d_esc8_pos.p, c->d_esc8_val.p,
                   (unsigned long long)col_cap_used, c->d_esc_pos.p, c->d_esc_val.p, (unsigned long long)cnt_cap);
        else if (gap_bytes == 1)
            LAUNCH(c, (pc::k_csr_pack8<256, unsigned char>), (unsigned int)c->n_rows, 256, c->d_indptr.p, c->d_indices.p, c->d_data.p, c->d_dcol8.p,
                   c->d_cnt8.p, c->d_counter.p + 8, c->d_esc8_pos.p, c->d_esc8_val.p, (unsigned long long)col_cap_used, c->d_esc_pos.p,
                   c->d_esc_val.p, (unsigned long long)cnt_cap);
        else
            LAUNCH(c, (pc::k_csr_pack8<256, unsigned short>), (unsigned int)c->n_rows, 256, c->d_indptr.p, c->d_indices.p, c->d_data.p,
                   reinterpret_cast<unsigned short*>(c->d_dcol8.p), c->d_cnt8.p, c->d_counter.p + 8, c->d_esc8_pos.p, c->d_esc8_val.p,
                   (unsigned long long)col_cap_used, c->d_esc_pos.p, c->d_esc_val.p, (unsigned long long)cnt_cap);
    }
    READ_SMALL(c, ne, c->d_counter.p + 8, sizeof ne);
    if (indptr) { rc = copy_out(c, indptr, c->d_indptr.p, (c->n_rows + 1) * sizeof(int64_t), loc); if (rc) return rc; }
    if (dcol8) { rc = copy_out(c, dcol8, c->d_dcol8.p, c->nnz * gap_bytes, loc); if (rc) return rc; }
    if (cnt8 && !g12c4) { rc = copy_out(c, cnt8, c->d_cnt8.p, c->nnz, loc); if (rc) return rc; }
    SYNC(c);
    if ((int64_t)ne[0] > col_cap_used || (int64_t)ne[1] > cnt_cap)
        return fail(PC_ESTATE, "escape lists of the 8-bit hand-off overflow (%llu columns, %llu counts): use pc_csr_get16", ne[0], ne[1]);
    c->n_esc8_col = (int64_t)ne[0]; c->n_esc8_cnt = (int64_t)ne[1]; c->n_esc = (int64_t)ne[1];
    n_escapes[0] = c->n_esc8_col; n_escapes[1] = c->n_esc8_cnt;
    return PC_OK;
}

int pc_csr_escapes8_get(pc_ctx* c, int64_t* col_pos, int32_t* col_val, int64_t* cnt_pos, float* cnt_val, int loc) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->built || c->n_esc8_col < 0) return fail(PC_ESTATE, "pc_csr_escapes8_get before pc_csr_get8");
    if (c->n_esc8_col > 0) {
        if (!col_pos || !col_val) return fail(PC_EINVAL, "null column escape buffers");
        rc = copy_out(c, col_pos, c->d_esc8_pos.p, c->n_esc8_col * sizeof(int64_t), loc); if (rc) return rc;
        rc = copy_out(c, col_val, c->d_esc8_val.p, c->n_esc8_col * sizeof(int32_t), loc); if (rc) return rc;
    }
    if (c->n_esc8_cnt > 0) {
        if (!cnt_pos || !cnt_val) return fail(PC_EINVAL, "null count escape buffers");
        rc = copy_out(c, cnt_pos, c->d_esc_pos.p, c->n_esc8_cnt * sizeof(int64_t), loc); if (rc) return rc;
        rc = copy_out(c, cnt_val, c->d_esc_val.p, c->n_esc8_cnt * sizeof(float), loc); if (rc) return rc;
    }
    return PC_OK;
}

int pc_csr_escapes_get(pc_ctx* c, int64_t* pos, float* val, int loc) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->built || c->n_esc < 0) return fail(PC_ESTATE, "pc_csr_escapes_get before pc_csr_get16");
    if (c->n_esc == 0) return PC_OK;
    if (!pos || !val) return fail(PC_EINVAL, "null escape buffers");
    rc = copy_out(c, pos, c->d_esc_pos.p, c->n_esc * sizeof(int64_t), loc); if (rc) return rc;
    return copy_out(c, val, c->d_esc_val.p, c->n_esc * sizeof(float), loc);
}

int pc_tfidf_factors_get(pc_ctx* c, float* idf, double* row_norm, int loc) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->tfidf_done) return fail(PC_ESTATE, "no tf-idf computed");
    StageTimer t(c, ST_D2H);
    if (idf) { rc = copy_out(c, idf, c->d_idf.p, c->n_sel * sizeof(float), loc); if (rc) return rc; }
    if (row_norm) { rc = copy_out(c, row_norm, c->d_row_norm.p, c->n_rows * sizeof(double), loc); if (rc) return rc; }
    return PC_OK;
}

int pc_df_get(pc_ctx* c, int32_t* df, int loc) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->built) return fail(PC_ESTATE, "no matrix built");
    return copy_out(c, df, c->d_df.p, c->n_sel * sizeof(int32_t), loc);
}

// ---------------------------------------------------------------------------------------------- K7
int pc_tfidf(pc_ctx* c, int64_t n_rows_global, const int32_t* df_global, int df_loc) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->built) return fail(PC_ESTATE, "pc_tfidf before pc_build_matrix");
    if (n_rows_global <= 0) n_rows_global = c->n_rows;
    c->tfidf_done = false;
    rc = c->d_idf.alloc(std::max<int64_t>(c->n_sel, 1)); if (rc) return rc;
    rc = c->d_tfidf.alloc(std::max<int64_t>(c->nnz, 1)); if (rc) return rc;
    const int* df = c->d_df.p;
    if (df_global) {
        if (df_loc == PC_DEVICE) df = df_global;
        else {
            rc = c->d_df_global.alloc(std::max<int64_t>(c->n_sel, 1)); if (rc) return rc;
            if (c->n_sel) CU(cudaMemcpyAsync(c->d_df_global.p, df_global, c->n_sel * sizeof(int), cudaMemcpyHostToDevice, c->stream));
            df = c->d_df_global.p;
        }
    }
    {
        StageTimer t(c, ST_TFIDF);
        if (c->n_sel > 0) LAUNCH(c, pc::k_idf, blocks_for(c->n_sel, 256), 256, df, c->n_sel, (float)(n_rows_global + 1), c->d_idf.p);
        rc = c->d_row_norm.alloc(std::max<int64_t>(c->n_rows, 1)); if (rc) return rc;
        CU(cudaMemsetAsync(c->d_row_norm.p, 0, std::max<int64_t>(c->n_rows, 1) * sizeof(double), c->stream));
        if (c->n_rows > 0 && c->nnz > 0)
            LAUNCH(c, (pc::k_tfidf<256>), (unsigned int)c->n_rows, 256, c->d_indptr.p, c->d_indices.p, c->d_data.p, c->d_idf.p, c->d_tfidf.p,
                   c->d_row_norm.p);
    }
    SYNC(c);
    c->tfidf_done = true;
    return PC_OK;
}

int pc_tfidf_get(pc_ctx* c, float* data, int loc) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->tfidf_done) return fail(PC_ESTATE, "no tf-idf computed");
    StageTimer t(c, ST_D2H);
    return copy_out(c, data, c->d_tfidf.p, c->nnz * sizeof(float), loc);
}

// ---------------------------------------------------------------------------------------------- N3 standard scaling
int pc_standard_scale(pc_ctx* c, const double* row_scale, int loc) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->built) return fail(PC_ESTATE, "pc_standard_scale before pc_build_matrix");
    c->scaled_done = false;
    const int64_t R = c->n_rows, D = std::max<int64_t>(c->n_sel, 1);
    rc = c->d_scaled.alloc(std::max<int64_t>(c->nnz, 1)); if (rc) return rc;
    rc = c->d_colstat.alloc((size_t)3 * D); if (rc) return rc;
    const double* rs = nullptr;
    if (row_scale) {
        if (loc == PC_DEVICE) rs = row_scale;
        else {
            rc = c->d_rowscale.alloc(std::max<int64_t>(R, 1)); if (rc) return rc;
            if (R) CU(cudaMemcpyAsync(c->d_rowscale.p, row_scale, (size_t)R * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            rs = c->d_rowscale.p;
        }
    }
    double* col_sum = c->d_colstat.p; double* col_ss = col_sum + D; double* factor = col_ss + D;
    CU(cudaMemsetAsync(c->d_colstat.p, 0, (size_t)3 * D * sizeof(double), c->stream));
    if (R > 0 && c->nnz > 0) {
        StageTimer t(c, ST_TFIDF);
        LAUNCH(c, (pc::k_col_sum<256>), (unsigned int)R, 256, c->d_indptr.p, c->d_indices.p, c->d_data.p, rs, col_sum);
        LAUNCH(c, (pc::k_col_sqdev<256>), (unsigned int)R, 256, c->d_indptr.p, c->d_indices.p, c->d_data.p, rs, col_sum, (double)R, col_ss);
        LAUNCH(c, pc::k_col_factor, blocks_for(c->n_sel, 256), 256, col_sum, col_ss, c->d_df.p, c->n_sel, (double)R, factor);
        LAUNCH(c, (pc::k_col_apply<256>), (unsigned int)R, 256, c->d_indptr.p, c->d_indices.p, c->d_data.p, rs, factor, c->d_scaled.p);
    }
    SYNC(c);
    c->scaled_done = true;
    return PC_OK;
}

int pc_scaled_get(pc_ctx* c, float* data, int loc) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->scaled_done) return fail(PC_ESTATE, "no standard-scaled values computed");
    return copy_out(c, data, c->d_scaled.p, c->nnz * sizeof(float), loc);
}

// ---------------------------------------------------------------------------------------------- N3 Gram matrix panels
int pc_densify_panel(pc_ctx* c, int which, int64_t col0, int ncols, float* out_dev, int64_t ld) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->built) return fail(PC_ESTATE, "pc_densify_panel before pc_build_matrix");
    if (col0 < 0 || ncols < 1 || ld < ncols || !out_dev) return fail(PC_EINVAL, "bad pc_densify_panel arguments");
    const float* values = nullptr;
    if (which == 0) values = c->d_data.p;
    else if (which == 1) { if (!c->tfidf_done) return fail(PC_ESTATE, "no tf-idf computed"); values = c->d_tfidf.p; }
    else if (which == 2) { if (!c->scaled_done) return fail(PC_ESTATE, "no standard-scaled values computed"); values = c->d_scaled.p; }
    else return fail(PC_EINVAL, "which must be 0 (counts), 1 (tf-idf) or 2 (standard-scaled)");
    if (c->n_rows > 0) {
        StageTimer t(c, ST_GRAM);
        LAUNCH(c, (pc::k_densify_panel<float, 256>), (unsigned int)c->n_rows, 256, c->d_indptr.p, c->d_indices.p, values, col0, ncols,
               out_dev, ld, c->gram_preround);
    }
    return PC_OK;
}

// K += P P^T on the tensor cores (gram_tc.cuh): P = panel_dev, n_rows x ncols float32, row-major with leading dimension
// ld (a multiple of 4 elements, 16-byte aligned base); K = k_dev, n_rows x n_rows float32, leading dimension ldk.  Only
// the tiles on and above the diagonal are updated; pc_gram_finish mirrors them.  The panel is rounded to TF32 in place.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else cudaGetLastError();
    }
    return fn;
}

int pc_gram_accumulate(pc_ctx* c, float* panel_dev, int64_t n_rows, int ncols, int64_t ld, float* k_dev, int64_t ldk) {
    int rc = use_device(c); if (rc) return rc;
    if (!panel_dev || !k_dev || n_rows < 1 || n_rows > 0x7FFFFFFF || ncols < 1 || ld < ncols || ldk < n_rows)
        return fail(PC_EINVAL, "bad pc_gram_accumulate arguments");
    if ((ld & 3) || (reinterpret_cast<uintptr_t>(panel_dev) & 15))
        return fail(PC_EINVAL, "the panel must be 16-byte aligned with a leading dimension that is a multiple of 4 floats (TMA)");
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return fail(PC_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)ncols, (cuuint64_t)n_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)pc::gram::BK, (cuuint32_t)pc::gram::BM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, panel_dev, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PC_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    const int nt = (int)((n_rows + pc::gram::BM - 1) / pc::gram::BM);
    const int n_kblocks = (ncols + pc::gram::BK - 1) / pc::gram::BK;
    StageTimer t(c, ST_GRAM);
    if (!c->gram_preround)                                     // (pc_tune gram_preround 1: pc_densify_panel already rounded the panel)
        LAUNCH(c, pc::gram::k_round_tf32, (unsigned int)std::min<int64_t>(148 * 16, (n_rows * ncols + 255) / 256), 256, panel_dev, n_rows, ncols, ld);
    if (c->gram_variant == 1) {
        // persistent 128 x 256 tiles, two TMEM accumulators (gram_tc.cuh, namespace v2): a second tensor map with a 256-row box
        CUtensorMap tm_b;
        const cuuint32_t box_b[2] = {(cuuint32_t)pc::gram::BK, (cuuint32_t)pc::gram::v2::BN2};
        r = encode(&tm_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, panel_dev, dims, strides, box_b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(PC_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
        const int ntn = (int)((n_rows + pc::gram::v2::BN2 - 1) / pc::gram::v2::BN2);
        int n_tiles = 0;
        for (int ti = 0; ti < nt; ++ti) n_tiles += ntn - (ti >> 1);
        CU(cudaFuncSetAttribute(pc::gram::v2::k_gram_tf32_p, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pc::gram::v2::SMEM_BYTES2));
        c->launches++;
        pc::gram::v2::k_gram_tf32_p<<<(unsigned int)std::min(n_tiles, 148), pc::gram::THREADS, pc::gram::v2::SMEM_BYTES2, c->stream>>>(
            tm, tm_b, k_dev, (int)n_rows, ldk, n_kblocks, ntn, n_tiles);
        CU(cudaGetLastError());
        return PC_OK;
    }
    CU(cudaFuncSetAttribute(pc::gram::k_gram_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pc::gram::SMEM_BYTES));
    c->launches++;
    pc::gram::k_gram_tf32<<<(unsigned int)(nt * (nt + 1) / 2), pc::gram::THREADS, pc::gram::SMEM_BYTES, c->stream>>>(
        tm, k_dev, (int)n_rows, ldk, n_kblocks, nt);
    CU(cudaGetLastError());
    return PC_OK;
}

int pc_gram_finish(pc_ctx* c, float* k_dev, int64_t n_rows, int64_t ldk) {
    int rc = use_device(c); if (rc) return rc;
    if (!k_dev || n_rows < 1 || n_rows > 0x7FFFFFFF || ldk < n_rows) return fail(PC_EINVAL, "bad pc_gram_finish arguments");
    LAUNCH(c, pc::gram::v2::k_gram_mirror_all, (unsigned int)std::min<int64_t>(148 * 16, (n_rows * n_rows + 255) / 256), 256, k_dev, (int)n_rows, ldk);
    return PC_OK;
}

// ---------------------------------------------------------------------------------------------- N2 diagnostics
int pc_count_histogram(pc_ctx* c, uint32_t min_count, uint32_t max_count, uint64_t* hist, int loc) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->counted) return fail(PC_ESTATE, "pc_count_histogram before pc_count");
    if (!hist || min_count < 1 || max_count < min_count || (uint64_t)max_count - min_count > (1u << 26))
        return fail(PC_EINVAL, "bad pc_count_histogram arguments");
    if (c->count_kind == 1 && min_count < c->list_min)
        return fail(PC_ESTATE, "k-mers with fewer than %u occurrences were not kept by this count (pc_tune keep_min); asked for >= %u",
                    c->list_min, min_count);
    const size_t bins = (size_t)(max_count - min_count) + 2;
    DevBuf<unsigned long long> d;
    rc = d.alloc(bins); if (rc) return rc;
    CU(cudaMemsetAsync(d.p, 0, bins * sizeof(unsigned long long), c->stream));
    if (c->count_kind == 1) {
        if (c->n_list > 0)
            LAUNCH(c, pc::k_count_hist_list, (unsigned int)std::min<int64_t>((c->n_list + 255) / 256, 148 * 8), 256,
                   c->list_counts.p, c->n_list, min_count, max_count, d.p);
    } else {
        LAUNCH(c, pc::k_count_hist_table, (unsigned int)std::min<uint64_t>((c->capacity + 255) / 256, 148ull * 16), 256,
               c->table(), c->capacity, min_count, max_count, d.p);
    }
    rc = copy_out(c, hist, d.p, bins * sizeof(uint64_t), loc);
    if (!rc && loc == PC_DEVICE) rc = sync_stream(c);           // d is released below
    d.release();
    return rc;
}

int pc_row_hits(pc_ctx* c, int64_t* hits_per_row, int loc) {
    int rc = use_device(c); if (rc) return rc;
    if (!c->built) return fail(PC_ESTATE, "pc_row_hits before pc_build_matrix");
    if (!hits_per_row) return fail(PC_EINVAL, "hits_per_row is null");
    if (c->n_rows == 0) return PC_OK;
    DevBuf<int64_t> d;
    rc = d.alloc((size_t)c->n_rows); if (rc) return rc;
    LAUNCH(c, pc::k_row_hits, blocks_for(c->n_rows, 256), 256, c->d_row_cursor.p, c->n_rows, d.p);
    rc = copy_out(c, hits_per_row, d.p, (size_t)c->n_rows * sizeof(int64_t), loc);
    if (!rc && loc == PC_DEVICE) rc = sync_stream(c);
    d.release();
    return rc;
}

// small device -> host read-back through the mailbox (callers: the multi-GPU driver's bucket sizes, counts, handles)
int pc_read_back(pc_ctx* c, void* dst_host, const void* src_dev, int64_t bytes) {
    int rc = use_device(c); if (rc) return rc;
    if (bytes < 0 || (bytes > 0 && (!dst_host || !src_dev))) return fail(PC_EINVAL, "bad pc_read_back arguments");
    READ_SMALL(c, dst_host, src_dev, (size_t)bytes);
    SYNC(c);
    return PC_OK;
}

// ---------------------------------------------------------------------------------------------- SM-driven copies
int pc_copy_async(pc_ctx* c, void* dst, const void* src, int64_t bytes, void* stream, int n_ctas) {
    int rc = use_device(c); if (rc) return rc;
    if (bytes < 0 || (bytes > 0 && (!dst || !src))) return fail(PC_EINVAL, "bad pc_copy_async arguments");
    if (bytes == 0) return PC_OK;
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    if (((uintptr_t)dst | (uintptr_t)src) & 15) {              // unaligned: leave it to the copy engine
        CU(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, st));
        return PC_OK;
    }
    if (n_ctas <= 0) n_ctas = 32;
    const int64_t n16 = bytes / 16;
    if (n16 > 0) {
        unsigned int grid = (unsigned int)std::min<int64_t>(n_ctas, (n16 + 255) / 256);
        pc::k_copy16<<<grid, 256, 0, st>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), n16);
        CU(cudaGetLastError());
    }
    if (bytes & 15) {
        pc::k_copy_tail<<<1, 32, 0, st>>>(static_cast<const uint8_t*>(src) + n16 * 16, static_cast<uint8_t*>(dst) + n16 * 16, (int)(bytes & 15));
        CU(cudaGetLastError());
    }
    return PC_OK;
}

// ---------------------------------------------------------------------------------------------- measurement
int pc_timings(pc_ctx* c, float ms[PC_N_STAGES]) {
    int rc = use_device(c); if (rc) return rc;
    SYNC(c);
    for (int i = 0; i < PC_N_STAGES; ++i) ms[i] = 0.f;
    for (int i = 0; i < ST_N; ++i)
        if (c->ev_used[i]) { float v = 0.f; if (cudaEventElapsedTime(&v, c->ev[i][0], c->ev[i][1]) == cudaSuccess) ms[i] = v; else cudaGetLastError(); }
    return PC_OK;
}

// Start and end of every timed stage of the last step, in ms after the earliest stage start (-1: stage not used).  What
// the per-stage durations cannot show -- the device idling between stages while the host reads a size back or Python
// returns from one C call and enters the next -- is the difference between consecutive entries.
int pc_stage_timeline(pc_ctx* c, float start_ms[PC_N_STAGES], float end_ms[PC_N_STAGES]) {
    int rc = use_device(c); if (rc) return rc;
    if (!start_ms || !end_ms) return fail(PC_EINVAL, "null output");
    SYNC(c);
    for (int i = 0; i < PC_N_STAGES; ++i) start_ms[i] = end_ms[i] = -1.f;
    int first = -1;
    for (int i = 0; i < ST_N; ++i) {
        if (!c->ev_used[i]) continue;
        if (first < 0) { first = i; continue; }
        float v = 0.f;
        if (cudaEventElapsedTime(&v, c->ev[first][0], c->ev[i][0]) == cudaSuccess && v < 0.f) first = i; else cudaGetLastError();
    }
    if (first < 0) return PC_OK;
    for (int i = 0; i < ST_N; ++i) {
        if (!c->ev_used[i]) continue;
        float a = 0.f, b = 0.f;
        if (cudaEventElapsedTime(&a, c->ev[first][0], c->ev[i][0]) != cudaSuccess) { cudaGetLastError(); continue; }
        if (cudaEventElapsedTime(&b, c->ev[first][0], c->ev[i][1]) != cudaSuccess) { cudaGetLastError(); continue; }
        start_ms[i] = a; end_ms[i] = b;
    }
    return PC_OK;
}

int pc_launch_count(pc_ctx* c, int64_t* launches) {
    if (!c || !launches) return fail(PC_EINVAL, "bad arguments");
    *launches = c->launches;
    return PC_OK;
}

int pc_reset_counters(pc_ctx* c) {
    if (!c) return fail(PC_EINVAL, "null context");
    c->launches = 0;
    for (int i = 0; i < ST_N; ++i) c->ev_used[i] = false;
    return PC_OK;
}

// ---------------------------------------------------------------------------------------------- host unit-test hooks
int pc_host_pack_word(const uint8_t* bases, int n, uint64_t* word, uint32_t* nmask) {
    if (!bases || n < 0 || n > 32 || !word || !nmask) return fail(PC_EINVAL, "bad pack_word arguments");
    uint8_t tmp[32];
    for (int i = 0; i < 32; ++i) tmp[i] = i < n ? bases[i] : (uint8_t)'N';
    uint64_t w = 0; uint32_t nm = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t v = (uint32_t)tmp[4 * i] | ((uint32_t)tmp[4 * i + 1] << 8) | ((uint32_t)tmp[4 * i + 2] << 16) | ((uint32_t)tmp[4 * i + 3] << 24);
        w |= (uint64_t)pc::pack4_codes(v) << (56 - 8 * i);
        nm |= pc::pack4_invalid(v) << (28 - 4 * i);
    }
    for (int j = 0; j < 32; ++j) if ((nm >> (31 - j)) & 1u) w &= ~(3ull << (62 - 2 * j));
    *word = w; *nmask = nm;
    return PC_OK;
}

// host emulation of k_pack's layout rules with the same SWAR routines (unit tests of the bit tricks)
int pc_host_pack_chunks(const uint8_t* ascii, const int64_t* chunk_begin, const int64_t* chunk_end, int64_t n_chunks,
                        uint64_t* words, uint32_t* nmask, int64_t* chunk_word0, int64_t cap_words, int64_t* n_words) {
    if (!chunk_begin || !chunk_end || !n_words) return fail(PC_EINVAL, "bad pack_chunks arguments");
    int64_t wsum = 0;
    for (int64_t c = 0; c < n_chunks; ++c) wsum += (chunk_end[c] - chunk_begin[c]) / 32 + 1;
    *n_words = wsum;
    if (!words) return PC_OK;
    if (cap_words < wsum + 1) return fail(PC_EINVAL, "need %lld words", (long long)(wsum + 1));
    int64_t w = 0;
    for (int64_t c = 0; c < n_chunks; ++c) {
        int64_t len = chunk_end[c] - chunk_begin[c];
        if (chunk_word0) chunk_word0[c] = w;
        for (int64_t off = 0; off < len / 32 * 32 + 32; off += 32, ++w) {
            int64_t rem = len - off;
            int n = (int)(rem < 0 ? 0 : (rem > 32 ? 32 : rem));
            int rc = pc_host_pack_word(ascii + chunk_begin[c] + off, n, &words[w], &nmask[w]);
            if (rc) return rc;
        }
    }
    if (chunk_word0) chunk_word0[n_chunks] = w;
    words[w] = 0; nmask[w] = 0xFFFFFFFFu;
    return PC_OK;
}

// host run of the device Roller over packed words: 32 (canonical k-mer, valid) pairs per word
int pc_host_extract(const uint64_t* words, const uint32_t* nmask, int64_t n_words, int k, uint64_t* keys, uint8_t* valid) {
    if (!words || !nmask || !keys || !valid || k < 1 || k > 32) return fail(PC_EINVAL, "bad extract arguments");
    for (int64_t w = 0; w < n_words; ++w) {
        pc::Roller r;
        r.init(words[w], words[w + 1], nmask[w], nmask[w + 1], k);     // word n_words is the guard word
        for (int b = 0; b < 32; ++b) {
            keys[w * 32 + b] = r.canonical();
            valid[w * 32 + b] = r.valid(b) ? 1 : 0;
            if (b < 31) r.step(b);
        }
    }
    return PC_OK;
}

// host run of the super-k-mer extraction (the SAME inline routines as k_skm_extract: SkmThread, skm_run_length, skm_dest,
// skm_record) over packed words, and of the record expansion (RecRoller): unit tests of the bit logic without a GPU
extern "C++" {
template <int W>
static int64_t host_skm_records(const uint64_t* words, const uint32_t* nmask, int64_t n_words, int k, int m, int nmax, int pb,
                                unsigned int nb2, uint64_t* recs, int64_t cap) {
    int64_t n_out = 0;
    for (int64_t w = 0; w < n_words; ++w) {
        if (nmask[w] == 0xFFFFFFFFu) continue;
        pc::SkmThread<W> th;
        th.run(words[w], words[w + 1], nmask[w], nmask[w + 1], k, m, nmax);      // word n_words is the guard word
        uint32_t rem = th.starts;
        while (rem) {
            const int b = __builtin_ctz(rem);
            rem &= rem - 1;
            const int n = pc::skm_run_length(th.vmask, th.starts, b);
            const uint32_t dest = pc::skm_dest(th.hmin[b], pb, nb2);
            uint64_t hi, lo;
            pc::skm_record(words[w], words[w + 1], b, n, k, dest, hi, lo);
            if (recs && n_out < cap) { recs[2 * n_out] = hi; recs[2 * n_out + 1] = lo; }
            ++n_out;
        }
    }
    return n_out;
}
}  // extern "C++"

int pc_host_skm_plan(int k, int64_t n_sub_total, int32_t* m_out) {
    if (!m_out) return fail(PC_EINVAL, "m_out is null");
    *m_out = pc::skm_choose_m(k, n_sub_total);
    return PC_OK;
}

int pc_host_skm_records(const uint64_t* words, const uint32_t* nmask, int64_t n_words, int k, int m, int nmax, int pb,
                        int nb2, uint64_t* recs, int64_t cap, int64_t* n_records) {
    if (!words || !nmask || !n_records || k < 1 || k > 32 || m < 1 || m > 16 || m > k || nmax < 1 || nmax + k - 1 > PC_SKM_BASES ||
        pb < 0 || pb > 12 || nb2 < 1 || nb2 > PC_SUB_MAX)
        return fail(PC_EINVAL, "bad pc_host_skm_records arguments");
    const int W = k - m + 1;
    int64_t n = -1;
    switch (W) {
#define SKM_CASE(W_) case W_: n = host_skm_records<W_>(words, nmask, n_words, k, m, nmax, pb, (unsigned int)nb2, recs, cap); break;
        SKM_CASE(4) SKM_CASE(6) SKM_CASE(8) SKM_CASE(10) SKM_CASE(12) SKM_CASE(14) SKM_CASE(16) SKM_CASE(18) SKM_CASE(20) SKM_CASE(22)
#undef SKM_CASE
        default: return fail(PC_EINVAL, "W = k - m + 1 = %d is not an even value in [4, 22]", W);
    }
    *n_records = n;
    return PC_OK;
}

int pc_host_skm_expand(const uint64_t* recs, int64_t n_records, int k, uint64_t* keys, uint32_t* dest, int64_t cap, int64_t* n_keys) {
    if (!recs || !n_keys || k < 1 || k > 32) return fail(PC_EINVAL, "bad pc_host_skm_expand arguments");
    int64_t out = 0;
    for (int64_t i = 0; i < n_records; ++i) {
        pc::RecRoller r;
        r.init(recs[2 * i], recs[2 * i + 1], k);
        const int n = pc::skm_rec_n(recs[2 * i + 1]);
        for (int j = 0; j < n; ++j) {
            if (keys && out < cap) { keys[out] = r.canonical(); if (dest) dest[out] = pc::skm_rec_dest(recs[2 * i + 1]); }
            ++out;
            r.step();
        }
    }
    *n_keys = out;
    return PC_OK;
}

int pc_host_canonical(uint64_t kmer, int k, uint64_t* out) {
    if (k < 1 || k > 32 || !out) return fail(PC_EINVAL, "bad canonical arguments");
    uint64_t x = kmer & pc::kmer_mask(k), rc = pc::revcomp_bits(x, k);
    *out = x < rc ? x : rc;
    return PC_OK;
}

}  // extern "C"
