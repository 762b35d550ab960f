/* polycracker_b200 -- C ABI of the B200-native polyCRACKER hot path.
 *
 * The reference (jlevy44/PolyCRACKER-Unofficial-Mirror) has no FFI: its hot path is a chain of click
 * sub-commands that shell out to kmercountexact.sh / jellyfish / bbmap.sh / bedtools and glue the results with
 * CPython.  Each entry point below replaces one of those stages; the comment on each cites
 * the reference lines it stands in for (paths relative to /root/reference/polycracker/).  The Python side
 * (polycracker_b200/binding.py) binds exactly these symbols with ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 (PC_OK) or a negative PC_E* code; pc_last_error() gives the message
 *     (the reference ignores tool failures, polycracker.py:204-207; this library fails loudly instead).
 *   - plain pointers and sizes only.  `loc` says where a caller-owned buffer lives: PC_HOST or PC_DEVICE
 *     (a torch tensor's data_ptr()).  Sizes are queried first (pc_*_info), buffers are caller-owned.
 *   - one pc_ctx per GPU, one host thread per pc_ctx.  All work is issued on the ctx stream (pc_set_stream lets a
 *     caller hand in torch's current stream).
 *   - k-mers are uint64: first base in the most significant of the low 2k bits, A=0 C=1 G=2 T=3, canonical =
 *     min(k-mer, reverse complement) (jellyfish -C orientation, polycracker.py:203).  Integer order == the
 *     Python string order the reference sorts its columns by (polycracker.py:295).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns PC_ECUDA.
 */
#ifndef POLYCRACKER_B200_H
#define POLYCRACKER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PC_OK        0
#define PC_EINVAL   -1   /* bad argument / unsupported mode (k > 32, perfect_mode = 0, ...) */
#define PC_ECUDA    -2   /* CUDA runtime error (including "no device") */
#define PC_ESTATE   -3   /* stage called out of order */
#define PC_ENOMEM   -4   /* table would not fit / probe limit hit */
#define PC_EIO      -5   /* file error (host helpers) */

#define PC_HOST      0
#define PC_DEVICE    1

typedef struct pc_ctx pc_ctx;

/* ---- library ------------------------------------------------------------------------------------------- */
int         pc_version(void);
const char* pc_last_error(void);
int         pc_device_count(void);

/* ---- context ------------------------------------------------------------------------------------------- */
int pc_create(int device, pc_ctx** out);
int pc_destroy(pc_ctx* ctx);
int pc_set_stream(pc_ctx* ctx, void* cuda_stream);       /* NULL -> the ctx's own stream */
int pc_sync(pc_ctx* ctx);

/* ---- a1  splitFasta geometry (polycracker.py:129-156), host-side integer work ------------------------------
 * For every sequence length L: windows [i*C,(i+1)*C) for consecutive pairs of arange(0,L,C) when L > C, then the
 * tail [last, L).  Zero-length sequences yield nothing.  Output order: input sequence order, start ascending
 * (the caller applies `bedtools sort` / Python sorted() to names).  pc_chunk_count returns the number of chunks. */
int64_t pc_chunk_count(const int64_t* seq_len, int64_t n_seqs, int64_t chunk_size);
int     pc_chunk_table(const int64_t* seq_len, int64_t n_seqs, int64_t chunk_size,
                       int64_t* chunk_seq, int64_t* chunk_start, int64_t* chunk_end);

/* ---- K1  2-bit packing + N-mask (replaces `bedtools getfasta` + FASTA parsing, polycracker.py:170) ----------
 * ascii: newline-free sequence bytes (n_bytes), host or device.  Chunk c is ascii[chunk_begin[c], chunk_end[c])
 * (host arrays, ascending, non-overlapping).  Anything but acgtACGT is masked.  Each chunk is an independent
 * record: no k-mer spans two chunks. */
int pc_load_sequences(pc_ctx* ctx, const uint8_t* ascii, int64_t n_bytes, int loc,
                      const int64_t* chunk_begin, const int64_t* chunk_end, int64_t n_chunks);
/* device-resident staging buffer (>= n_bytes) a caller may fill itself before pc_load_sequences(ascii=that ptr,
 * loc=PC_DEVICE); avoids a copy when the genome is produced on the GPU. */
int pc_ascii_buffer(pc_ctx* ctx, int64_t n_bytes, uint8_t** dev_ptr);
/* copy the packed representation out (tests): words[n_words] uint64, nmask[n_words] uint32, chunk_word0[n_chunks+1] */
int pc_packed_info(pc_ctx* ctx, int64_t* n_words, int64_t* n_chunks);
int pc_packed_get(pc_ctx* ctx, uint64_t* words, uint32_t* nmask, int64_t* chunk_word0, int loc);

/* ---- K2+K3  canonical k-mer counting (replaces kmercountexact.sh / jellyfish count -C, polycracker.py:199,203)
 * capacity = 0 -> sized from the number of windows.  1 <= k <= 32. */
int pc_count(pc_ctx* ctx, int k, int64_t capacity);
/* info[0]=capacity info[1]=distinct k-mers info[2]=valid windows info[3]=table bytes info[4]=max probe length */
int pc_count_info(pc_ctx* ctx, int64_t info[8]);
/* how the last count is held.  detail[0] = 0: global open-addressing table; 1: compact (k-mer, count) list produced by
 * the shared-memory count (two-level radix partition, one shared-memory table per sub-bucket; pc_tune part_mode 2 or
 * auto).  For the list: detail[1] = entries, detail[2] = lowest count kept (pc_tune keep_min: kmercountexact only ever
 * dumps mincount=3, polycracker.py:199, so k-mers below it never have to leave the SM), detail[3] = sub-buckets,
 * detail[4] = sub-buckets per level-1 bucket, detail[5] = sub-buckets re-counted through a global table because their
 * distinct k-mers did not fit the shared-memory table; super-k-mer count: detail[6] = records, detail[7] = minimizer
 * length m (both 0 when the k-mers themselves were partitioned). */
int pc_count_detail(pc_ctx* ctx, int64_t detail[8]);
/* Counters are uint32.  A counting unit (one shared-memory table, one sub-table, the direct table) that received >= 2^32
 * k-mer instances could wrap one (one k-mer with > 4.29e9 occurrences on one GPU): such a count is REFUSED with PC_ENOMEM,
 * never returned wrapped (SURVEY 8c "counts saturate / flag at 2^32 - 1").  flags[0] = 1 if the last count could have
 * wrapped (always 0 for a count that was returned), flags[1] = k-mer instances of the largest counting unit (upper bound for
 * the super-k-mer count), flags[2] = sub-buckets that overflowed their shared-memory table and were re-counted through a
 * global table, flags[3] = slots per table. */
int pc_count_flags(pc_ctx* ctx, int64_t flags[4]);
/* dump (kmer, count) with min_count <= count <= max_count (the .kcount file: `mincount=3` for k<=31,
 * polycracker.py:199).  keys == NULL -> only *n is written.  Unordered, like the tools' dumps. */
int pc_dump_counts(pc_ctx* ctx, uint32_t min_count, uint32_t max_count,
                   uint64_t* keys, uint32_t* counts, int64_t cap, int loc, int64_t* n);
/* multi-GPU counting.  Bucket = top pb bits of the k-mer hash; rank r owns a contiguous bucket range.
 * pc_partition: K1 must have run; extracts the canonical k-mers of the local chunks and radix-partitions them into
 * 2^pb buckets.  bucket_off[2^pb + 1] (host) receives the bucket offsets into *keys_dev (device, bucket-sorted k-mers).
 * The slices other ranks own travel as raw 8-byte k-mers in one all-to-all (torch.distributed / NCCL, caller side).
 * pc_count_buckets: counts n_buckets buckets into a fresh table; bucket i consists of the n_seg segments
 * keys_dev[seg_begin[i*n_seg+s] .. +seg_len[i*n_seg+s]) (one segment per sending rank).  Afterwards pc_select /
 * pc_dump_counts see the owner's share of the global counts. */
int pc_partition(pc_ctx* ctx, int k, int pb, int64_t* bucket_off, uint64_t** keys_dev);
int pc_count_buckets(pc_ctx* ctx, int k, int pb, int n_buckets, const uint64_t* keys_dev, const int64_t* seg_begin,
                     const int64_t* seg_len, int n_seg, int64_t capacity);
/* Peer-memory variant (one node, NVLink): instead of an all-to-all into a receive buffer the owner's count kernel
 * streams each sender's slice straight out of that sender's bucket buffer.  pc_kbuf_export pins the bucket buffer at
 * >= n_keys k-mers and returns its 64-byte CUDA IPC handle; pc_peer_open maps a peer's handle into this process;
 * pc_count_buckets_from is pc_count_buckets with one base pointer per segment (host array of device pointers,
 * seg_begin = offset inside that sender's buffer).  The caller orders partition -> count -> next partition across
 * ranks (the bucket-size all-gather before and the repeat-set all-gather after the count already do). */
int pc_kbuf_export(pc_ctx* ctx, int64_t n_keys, void* ipc_handle_64);
int pc_peer_open(pc_ctx* ctx, const void* ipc_handle_64, uint64_t** dev_ptr);
int pc_peer_close(pc_ctx* ctx, uint64_t* dev_ptr);
int pc_count_buckets_from(pc_ctx* ctx, int k, int pb, int n_buckets, const uint64_t* const* seg_base,
                          const int64_t* seg_begin, const int64_t* seg_len, int n_seg, int64_t capacity);
/* Shared-memory count on several GPUs: the owner splits each of its buckets into nb2 sub-buckets and needs their sizes
 * before it scatters.  Computing them itself would stream every peer segment over NVLink twice; instead every sender
 * histograms its own bucket buffer (local HBM) and the small histograms are exchanged.
 * pc_sub_plan: sub-buckets per bucket the count would choose for `total` k-mers in n_buckets buckets (0 when the
 * shared-memory count is not in use: skip the rest).  pc_sub_hist: after pc_partition, hist_dev[q * nb2 + x] (device,
 * uint64, 2^pb * nb2 entries) = local k-mers of bucket q in sub-bucket x.  pc_sub_hist_set: stage the owner's summed
 * histogram (device, uint64[n = n_buckets * nb2], must stay valid until the count returns) for the NEXT
 * pc_count_buckets / pc_count_buckets_from call, which then skips its own histogram pass. */
int pc_sub_plan(pc_ctx* ctx, int64_t total, int n_buckets);
int pc_sub_hist(pc_ctx* ctx, int pb, int nb2, uint64_t* hist_dev);
int pc_sub_hist_set(pc_ctx* ctx, int nb2, const uint64_t* hist_dev, int64_t n);
/* Super-k-mer variant of the multi-GPU count (pc_tune part_mode 3 / auto for k >= 17; kernels_skm.cuh).  What travels is
 * not one 8-byte k-mer per window but 16-byte records: a run of consecutive windows that share their minimizer, as
 * 2-bit packed bases + destination + length.  All instances of a canonical k-mer share the minimizer, so the owner of a
 * bucket still sees every instance of its k-mers and the counts stay exact (same role as kmercountexact / jellyfish,
 * polycracker.py:199, 203).
 * pc_skm_plan: plan shared by all ranks from the GLOBAL number of windows: plan[0] = pb, [1] = sub-buckets per level-1
 *   bucket, [2] = minimizer length m, [3] = W = k - m + 1, [4] = k-mers per record at most, [5] = 1 when usable for this k.
 * pc_skm_partition: K1 must have run.  bucket_off[2^pb + 1] (host) = level-1 bucket offsets in RECORDS into *recs_dev
 *   (device, 16 bytes per record); *hist_dev (device, uint64[2^pb * nb2]) = (k-mers << 32) | records of every sub-bucket.
 * pc_skm_export: after pc_load_sequences; sizes the record buffer for the loaded chunks (it only ever grows, so the handle
 *   stays valid from step to step) and returns its CUDA IPC handle (peers open it with pc_peer_open).  pc_skm_count_from: counts n_buckets buckets given as record segments (seg_base[i * n_seg + s] device
 *   pointers, or all relative to recs_dev when seg_base is NULL; begin / length in records) with hist_dev = the packed
 *   histogram of exactly these buckets summed over the senders (device, uint64[n_buckets * nb2]). */
int pc_skm_plan(pc_ctx* ctx, int k, int64_t windows_global, int world, int32_t plan[8]);
int pc_skm_partition(pc_ctx* ctx, int k, const int32_t plan[8], int64_t* bucket_off, uint64_t** recs_dev, uint64_t** hist_dev);
int pc_skm_export(pc_ctx* ctx, int k, const int32_t plan[8], void* ipc_handle_64);
int pc_skm_count_from(pc_ctx* ctx, int k, const int32_t plan[8], int n_buckets, const uint64_t* recs_dev,
                      const uint64_t* const* seg_base, const int64_t* seg_begin, const int64_t* seg_len, int n_seg,
                      const uint64_t* hist_dev);
/* alternative exchange of pre-aggregated pairs: all (kmer,count) pairs of the local table grouped by owner rank
 * (owner = hash(kmer) % world); offsets[world+1] is a host array.  keys/counts are device buffers of
 * capacity >= distinct k-mers.  pc_merge_counts adds received pairs into the ctx table (creating it with
 * `capacity` slots when *fresh* != 0). */
int pc_export_by_owner(pc_ctx* ctx, int world, uint64_t* keys_dev, uint32_t* counts_dev, int64_t cap,
                       int64_t* offsets);
int pc_merge_counts(pc_ctx* ctx, const uint64_t* keys_dev, const uint32_t* counts_dev, int64_t n,
                    int k, int64_t capacity, int fresh);

/* ---- N4  externally counted k-mers (scope row N4, second half; polycracker/format.py:123-133 hipmer_output_to_kcount and
 * any other .kcount the reference's kmer2Fasta would read, polycracker.py:220-238).  The (k-mer, count) pairs of such a file
 * become the count result of the context, exactly as if pc_count had produced them, so pc_select / pc_dump_counts /
 * pc_count_histogram and everything after them (K4..K7) run unchanged on k-mers the GPU never counted.  Keys are
 * strand-collapsed on entry; a k-mer named twice (or by both strands) and an entry that is no k-mer are refused. */
int pc_load_counts(pc_ctx* ctx, const uint64_t* keys, const uint32_t* counts, int64_t n, int k, int loc);

/* ---- K4  repeat-k-mer selection (replaces kmer2Fasta, polycracker.py:220-238) ------------------------------
 * keep count >= low (and <= high when use_high); sort ascending; keep every `sampling`-th survivor
 * (sorted-order sampling: declared deviation, the reference samples in the counter's dump order). */
int pc_select(pc_ctx* ctx, uint32_t low, uint32_t high, int use_high, int64_t sampling, int64_t* n_selected);
/* optional, before pc_count / pc_count_buckets: fuse K4 into K3.  The insert that lifts a k-mer to exactly `low`
 * occurrences records it, so a following pc_select(low, use_high = 0) does not re-scan the table (12 B x capacity of
 * HBM traffic).  low = 0 switches the watch off.  Any other pc_select argument falls back to the scan. */
int pc_watch_threshold(pc_ctx* ctx, uint32_t low);
int pc_selected_get(pc_ctx* ctx, uint64_t* keys, int loc);             /* n_selected keys, ascending */
/* install an externally chosen repeat set (the separately invoked blast_kmers stage reads it from the k-mer
 * FASTA, polycracker.py:244; the multi-GPU path installs the all-gathered set).  Keys need not be sorted or
 * canonical-unique on entry; they are canonicalised, sorted and de-duplicated; *n_unique is returned. */
int pc_set_selected(pc_ctx* ctx, const uint64_t* keys, int64_t n, int k, int loc, int64_t* n_unique);

/* ---- K5+K6  exact matching + occurrence matrix (replaces bbmap.sh perfectmode=t, blast2bed's sort/merge and
 * generate_Kmer_Matrix's Counter/dok/tocsc, polycracker.py:252, 268-280, 343-355) --------------------------
 * chunk_row[c] = row index of chunk c in the matrix or -1 when the chunk is filtered out
 * (polycracker.py:322-327).  Result: CSR (n_rows x n_selected), entry = number of windows of chunk i whose
 * canonical k-mer is repeat k-mer j.  Column indices ascending within a row. */
int pc_build_matrix(pc_ctx* ctx, const int32_t* chunk_row, int64_t n_rows, int64_t* nnz);
/* K6 alone, from hit lists that already exist (generate_Kmer_Matrix invoked on a blasted_merged.bed text artefact,
 * polycracker.py:343-355): row r holds column ids cols[row_offsets[r] .. row_offsets[r+1]) with multiplicity (host
 * arrays).  The Counter -> dok -> sparse work (sort + run-length encode per row, df) runs on the GPU. */
int pc_matrix_from_hits(pc_ctx* ctx, const int64_t* row_offsets, const int32_t* cols, int64_t n_rows, int64_t n_cols,
                        int64_t* nnz);
int pc_csr_get(pc_ctx* ctx, int64_t* indptr, int32_t* indices, float* data, int loc);
int pc_df_get(pc_ctx* ctx, int32_t* df, int loc);                      /* stored entries per column */
/* Compact hand-off of the same matrix (what crosses PCIe in the end-to-end path): the counts are integers bounded by the
 * windows of a chunk, so they travel as uint16 (6 bytes per entry with the int32 column id instead of 12 with the float32
 * count and tf-idf arrays).  An entry >= 65535 (one k-mer filling a chunk of >= 65535 windows: poly-A) is stored as 65535
 * and returned by pc_csr_escapes_get as (position in the CSR arrays, value); *n_escapes says how many there are.
 * The float32 arrays a consumer of clusteringMatrix.npz wants are rebuilt on the host (data = float32(data16), escapes
 * patched in). */
int pc_csr_get16(pc_ctx* ctx, int64_t* indptr, int32_t* indices, uint16_t* data16, int loc, int64_t* n_escapes);
int pc_csr_escapes_get(pc_ctx* ctx, int64_t* pos, float* val, int loc);
/* Densest hand-off, 2 bytes per entry: dcol8[i] = column(i) - column(i - 1) inside a row and cnt8[i] = count, one byte each;
 * 255 marks an escape -- the first entry of every row and every gap >= 255 carry their ABSOLUTE column in the column escape
 * list, counts >= 255 their value in the count escape list (positions index the CSR arrays; lists are unordered).
 * n_escapes[0] / [1] = entries of the two lists, fetched with pc_csr_escapes8_get.  pipeline.rebuild_indices / rebuild_counts
 * restore the int32 / float32 arrays on the host. */
int pc_csr_get8(pc_ctx* ctx, int64_t* indptr, uint8_t* dcol8, uint8_t* cnt8, int loc, int64_t n_escapes[2]);
/* The same hand-off with TWO-byte column gaps (escape 65535): for a large column universe (mean gap of a row n_selected /
 * entries per row beyond ~100: the 8-GPU weak-scaling run has 257) half of the one-byte gaps would escape at 12 bytes each;
 * 3 bytes per entry instead.  Escapes are fetched with pc_csr_escapes8_get as for pc_csr_get8. */
int pc_csr_get_gap16(pc_ctx* ctx, int64_t* indptr, uint16_t* dcol16, uint8_t* cnt8, int loc, int64_t n_escapes[2]);
/* ... and with gap and count sharing one uint16 ((gap << 4) | count; gap >= 4095 / count >= 15 escape as above): 2 bytes per
 * entry for a large column universe too.  What the end-to-end path ships when pc gap_dtype picks the wide form. */
int pc_csr_get_g12c4(pc_ctx* ctx, int64_t* indptr, uint16_t* packed, int loc, int64_t n_escapes[2]);
int pc_csr_escapes8_get(pc_ctx* ctx, int64_t* col_pos, int32_t* col_val, int64_t* cnt_pos, float* cnt_val, int loc);

/* ---- K7  TF-IDF + L2 rows (replaces sklearn TfidfTransformer().fit_transform, polycracker.py:395-396) -------
 * idf_j = ln((1+n)/(1+df_j)) + 1 in float32; rows divided by their L2 norm (double accumulation), zero rows
 * untouched.  n_rows_global / df_global (device or host, may be NULL = use local) let ranks share a global idf. */
int pc_tfidf(pc_ctx* ctx, int64_t n_rows_global, const int32_t* df_global, int df_loc);
int pc_tfidf_get(pc_ctx* ctx, float* data, int loc);                   /* nnz values, CSR order */
/* the two vectors the tf-idf values are a function of: idf[n_selected] (float32) and the L2 norm of every row (double,
 * 0 for an empty row).  value(i in row r, column j) = float32(double(float32(count_i) * idf_j) / row_norm_r) -- the
 * kernel's own arithmetic (one float32 multiply, one double divide rounded once), so a host rebuilds the array bit for
 * bit from 4 * n_selected + 8 * n_rows bytes instead of reading 4 * nnz. */
int pc_tfidf_factors_get(pc_ctx* ctx, float* idf, double* row_norm, int loc);

/* ---- N1  subgenome extraction inner loop (scope row N1; replaces the external-tool chain of
 * subgenomeExtraction, polycracker.py:991-1020: kmercountexact per subgenome 692-719, compareKmers 735-783,
 * bbmap map-back 786-799, blast2bed3 + bedtools coverage 802-865) ------------------------------------------
 * A session holds up to PC_MAX_GROUPS subgenomes and PC_MAX_SUB_K k-mer lengths.  Per length k:
 *   for each subgenome g: pc_load_sequences(its chunks); pc_count(k); pc_sub_add_counts(g, min_count)
 *   pc_sub_compare(...)   -> the differential k-mers of every subgenome for this k
 * then pc_load_sequences(all chunks); pc_sub_mapback(coverage).
 * min_count is the counter's dump floor (kmercountexact mincount=3 for k <= 31, jellyfish 1; pc.py:711-714).
 * compare: k-mer of g is differential if for ANY other subgenome o  count_g / count_o > ratio_threshold, where the
 * division floors when o holds the k-mer (Python 2 int / int, pc.py:765,778) and is count_g / default_value
 * otherwise.  sampling > 1 keeps every sampling-th differential k-mer in ascending k-mer order (the reference samples
 * in Python-2 dict order, which is not reproducible).  n_diff[n_groups] receives the set sizes.
 * diff_get: the set of (k, group), ascending, with counts[i * n_groups + o] = count in subgenome o (0 = absent).
 * mapback: coverage[chunk * n_groups + g] = number of distinct window starts in the chunk (loaded order) where a
 * differential k-mer of g, of any compared length, matches on either strand = column 6 of `bedtools coverage`. */
#define PC_MAX_GROUPS 8
#define PC_MAX_SUB_K  8
int pc_sub_begin(pc_ctx* ctx, int n_groups);
int pc_sub_add_counts(pc_ctx* ctx, int group, uint32_t min_count, int64_t* n_kmers);
int pc_sub_compare(pc_ctx* ctx, double ratio_threshold, double default_value, int64_t sampling, int64_t* n_diff);
int pc_sub_diff_get(pc_ctx* ctx, int k, int group, uint64_t* keys, uint32_t* counts, int loc, int64_t* n);
/* ---- N4  k-mer x genome count matrix (scope row N4; replaces the kmercountexact / dict / LabelEncoder / dok_matrix chain
 * of `bio_hyp_class`, utilities.py:177-284) -- in the same session as N1: after pc_sub_add_counts for every genome (one k),
 * pc_sub_union tests every k-mer of the UNION of the dictionaries: kept when max / clip(min, default_value) >=
 * ratio_threshold (integer division, Python 2 int arrays; min is 0 when a genome lacks the k-mer) and max >= max_val
 * (utilities.py:255-258).  *n_union = size of the union (rows of the reference's matrix before the filter), *n_kept = rows
 * kept.  pc_sub_union_get: the kept k-mers ascending (= LabelEncoder order for one k) and counts[i * n_groups + g]
 * (0 = absent) = union_kmer_matrix.npz / kmers_union.p.  The dictionaries stay pending (pc_sub_compare may follow). */
int pc_sub_union(pc_ctx* ctx, int64_t ratio_threshold, int64_t default_value, int64_t max_val, int64_t* n_union, int64_t* n_kept);
int pc_sub_union_get(pc_ctx* ctx, uint64_t* keys, uint32_t* counts, int loc, int64_t* n);
int pc_sub_mapback(pc_ctx* ctx, int32_t* coverage, int loc);

/* ---- N3  standard scaling (scope row N3, first half) --------------------------------------------------------
 * The normalisation the reference applies when tfidf = 0: generate_Kmer_Matrix's length normalisation
 * (polycracker.py:357-359: row i times row_scale[i] = 1 / (len_i / 5000.), value = float32(float64(v) * row_scale[i]);
 * row_scale NULL = raw counts) followed by sklearn StandardScaler(with_mean=False).fit_transform
 * (polycracker.py:392-393): every column divided by its standard deviation over all rows, near-constant columns
 * untouched.  Values in CSR order (same indptr / indices as pc_csr_get), float32, within 1e-6 relative of sklearn.
 * Single context only (the column statistics are not reduced across ranks). */
int pc_standard_scale(pc_ctx* ctx, const double* row_scale, int loc);
int pc_scaled_get(pc_ctx* ctx, float* data, int loc);

/* ---- N3, second half: the Gram matrix X X^T that KernelPCA(kernel = 'linear' | 'cosine') forms from the matrix rows
 * (polycracker.py:387 and 398-399: sklearn pairwise linear_kernel / cosine_similarity on the TF-IDF or standard-scaled rows).
 * The contraction runs over column panels densified on the fly: pc_densify_panel writes rows [0, n_rows) x columns
 * [col0, col0 + ncols) of the built matrix (which = 0 raw counts, 1 TF-IDF, 2 standard-scaled) as dense float32 into
 * out_dev (device, row-major, leading dimension ld >= ncols), asynchronously on the ctx stream.  polycracker_b200/gram.py
 * accumulates G += P P^T over the panels (pc_gram_accumulate below; or, as a cross-check, a plain cuBLAS GEMM through torch) and
 * applies the cosine normalisation.  pc_tune gram_preround 1 makes this call round the values to nearest TF32 for the tensor cores. */
int pc_densify_panel(pc_ctx* ctx, int which, int64_t col0, int ncols, float* out_dev, int64_t ld);
/* The contraction itself on the 5th-generation tensor cores (csrc/gram_tc.cuh: TMA-fed tcgen05.mma kind::tf32, TMEM
 * accumulators): K += P P^T for the tiles on and above the diagonal, P = panel_dev (n_rows x ncols, leading dimension ld a
 * multiple of 4 floats, 16-byte aligned; rounded to TF32 in place), K = k_dev (n_rows x n_rows float32, leading dimension
 * ldk, zeroed by the caller before the first panel).  pc_gram_finish mirrors the upper triangle once all panels are in.
 * Replaces the X X^T inside sklearn's linear_kernel / cosine_similarity (polycracker.py:387, 398-399); <= 1e-3 relative. */
int pc_gram_accumulate(pc_ctx* ctx, float* panel_dev, int64_t n_rows, int ncols, int64_t ld, float* k_dev, int64_t ldk);
int pc_gram_finish(pc_ctx* ctx, float* k_dev, int64_t n_rows, int64_t ldk);

/* ---- N2  diagnostics (scope row N2) ------------------------------------------------------------------------
 * pc_count_histogram: the data of `kcount_hist` (polycracker.py:1419-1496: Counter over the second column of a
 * .kcount): hist[c - min_count] = number of distinct k-mers that occur exactly c times, min_count <= c <= max_count;
 * hist[max_count - min_count + 1] = k-mers above max_count.  max_count - min_count + 2 entries (uint64, host or device).
 * pc_row_hits: the data of `number_repeatmers_per_subsequence` (polycracker.py:1366-1382: k-mer names per
 * blasted_merged.bed line): occurrences of repeat k-mers in every matrix row (rows with 0 have no line in the
 * reference's file), after pc_build_matrix / pc_matrix_from_hits. */
int pc_count_histogram(pc_ctx* ctx, uint32_t min_count, uint32_t max_count, uint64_t* hist, int loc);
int pc_row_hits(pc_ctx* ctx, int64_t* hits_per_row, int loc);

/* ---- memory ---------------------------------------------------------------------------------------------
 * Release device buffers later stages do not need (bit mask).  At Gbp scale the count table, the bucket buffer, the
 * k-mer stream and the hit buffers are tens of GB each: releasing the bucket buffer after pc_count and the table after
 * pc_select lets a 4.5 Gbp genome run on one 180 GB B200.  Releasing the stream makes K5 fall back to the roller. */
#define PC_REL_ASCII   1
#define PC_REL_BUCKETS 2
#define PC_REL_TABLE   4
#define PC_REL_STREAM  8
#define PC_REL_HITS    16
int pc_release(pc_ctx* ctx, int what);

/* Small device -> host read-back (bytes a multiple of 4, up to ~256 KB; larger or odd sizes fall back to a plain copy),
 * synchronous on the ctx stream.  Unlike cudaMemcpy it does not queue on a copy engine, so it is not delayed by a bulk
 * transfer of a neighbouring pipeline step (the context's own counters travel the same way). */
int pc_read_back(pc_ctx* ctx, void* dst_host, const void* src_dev, int64_t bytes);

/* ---- copies made by SMs -----------------------------------------------------------------------------------
 * dst/src: device memory or PAGE-LOCKED host memory (mapped into the device under unified addressing), 16-byte aligned
 * (otherwise the call falls back to cudaMemcpyAsync).  Enqueues a small grid (n_ctas CTAs, 0 = 32) on `stream` (NULL =
 * the ctx stream) that moves the bytes with 128-bit loads/stores.  For overlapping the bulk H2D / D2H of neighbouring
 * steps with this context's kernels without putting 400 MB transfers in front of its own tiny copies on the copy
 * engines (polycracker_b200/overlap.py). */
int pc_copy_async(pc_ctx* ctx, void* dst, const void* src, int64_t bytes, void* stream, int n_ctas);

/* ---- measurement ------------------------------------------------------------------------------------------
 * CUDA-event time (ms) of the last run of each stage on the ctx stream.  Index: 0 h2d, 1 pack, 2 table-init,
 * 3 count, 4 select(scan), 5 sort, 6 repeat-table, 7 probe, 8 row-sort, 9 emit, 10 tfidf, 11 d2h, 12 export, 13 merge,
 * 14 bucket histogram, 15 bucket scatter, 16 subgenome dictionary, 17 subgenome compare, 18 subgenome map-back,
 * 19 level-2 histogram + scatter of the shared-memory count, 20 global-table re-count of the sub-buckets that overflowed their
 * shared-memory table, 21 Gram matrix, 22 expansion of super-k-mer records into k-mers.
 * launches: kernels launched by this ctx since the last pc_reset_counters. */
#define PC_N_STAGES 24
int pc_timings(pc_ctx* ctx, float ms[PC_N_STAGES]);
/* start / end of every timed stage of the last step, ms after the earliest stage start (-1 = stage not used): shows the
 * device idle time between stages (host read-backs, the caller's own code between two calls) */
int pc_stage_timeline(pc_ctx* ctx, float start_ms[PC_N_STAGES], float end_ms[PC_N_STAGES]);
int pc_launch_count(pc_ctx* ctx, int64_t* launches);
int pc_reset_counters(pc_ctx* ctx);
/* kernel tunables (measurement / tuning only; defaults are the shipped configuration):
 * count_batch {4,8,16}, count_minb {2,3,4}, count_mode {0 load-first, 1 prefetch-ahead, 2 CAS-first} (direct table);
 * part_mode {-1 auto, 0 direct global table, 1 L2-partitioned global table, 2 shared-memory count}, part_mbytes,
 * part_blocks, part_casfirst, count_items {2,4}; keep_min (shared-memory count: lowest count kept, default 1),
 * smem_target (mean k-mers per sub-bucket, 0 = half a table), smem_variant {0..4: slots x threads x k-mers in flight},
 * smem_pb (level-1 bucket bits, 0 auto), scatter_tile {0 auto, 4096, 8192}, scatter_pack {0, 1 level 1, 2 both levels},
 * scatter_variant {0, 1 single returning atomic}; probe_filter {-1 auto, 0 off, 1 shared-memory bitmap, 2 L2 bitmap},
 * filter_bits, probe_threads, probe_fp, rowsort_warps {8, 16};
 * probe_mode {-1 auto, 0 roller, 1 stream}, probe_batch {4,8}, rep_load_inv; table_load_pct (5..95). */
int pc_tune(pc_ctx* ctx, const char* key, int64_t value);

/* ---- host helpers (no GPU) -------------------------------------------------------------------------------- */
/* pack one 32-base word on the host with the exact device routine (unit tests) */
int pc_host_pack_word(const uint8_t* bases, int n, uint64_t* word, uint32_t* nmask);
int pc_host_canonical(uint64_t kmer, int k, uint64_t* out);
/* host emulation of K1's layout rules and of the device k-mer roller (unit tests; same inline routines) */
int pc_host_pack_chunks(const uint8_t* ascii, const int64_t* chunk_begin, const int64_t* chunk_end, int64_t n_chunks,
                        uint64_t* words, uint32_t* nmask, int64_t* chunk_word0, int64_t cap_words, int64_t* n_words);
int pc_host_extract(const uint64_t* words, const uint32_t* nmask, int64_t n_words, int k, uint64_t* keys,
                    uint8_t* valid);
/* host run of the super-k-mer extraction and expansion with the device's own inline routines (unit tests):
 * records (two uint64 each: hi, lo) of the packed words; canonical k-mers (+ destination) of records, in record order */
int pc_host_skm_plan(int k, int64_t n_sub_total, int32_t* m_out);
int pc_host_skm_records(const uint64_t* words, const uint32_t* nmask, int64_t n_words, int k, int m, int nmax, int pb,
                        int nb2, uint64_t* recs, int64_t cap, int64_t* n_records);
int pc_host_skm_expand(const uint64_t* recs, int64_t n_records, int k, uint64_t* keys, uint32_t* dest, int64_t cap,
                       int64_t* n_keys);
/* kmer <-> text */
int pc_kmer_decode(const uint64_t* keys, int64_t n, int k, char* out /* n*k bytes, no separators */);
int pc_kmer_encode(const char* text, int64_t n, int k, uint64_t* keys /* ~0 when not acgtACGT */);
/* FASTA reader: two-call protocol.  pc_fasta_scan fills n_seqs / total bases / name bytes; pc_fasta_read fills
 * seq (newline-free), offsets[n_seqs+1], names ('\n'-joined first tokens of the headers). */
int pc_fasta_scan(const char* path, int64_t* n_seqs, int64_t* n_bases, int64_t* name_bytes);
int pc_fasta_read(const char* path, uint8_t* seq, int64_t* offsets, char* names);
/* text writers: "KMER\tCOUNT\n" (.kcount) and ">KMER\nKMER\n" (.kcount.fa) */
int pc_write_kcount(const char* path, const uint64_t* keys, const uint32_t* counts, int64_t n, int k, int append);
int pc_write_kmer_fasta(const char* path, const uint64_t* keys, int64_t n, int k, int rc_labels);

/* text artefacts of the mapping stages from a CSR (row = chunk, column = repeat k-mer): mode 0 headerless SAM
 * (polycracker.py:252; QNAME/RNAME as blast2bed reads them, POS 0), mode 1 blasted.bed, mode 2 blasted_merged.bed
 * (polycracker.py:272, 280).  names_blob/name_off: row labels back to back; row_len: chunk lengths. */
int pc_write_hit_lines(const char* path, int mode, const int64_t* indptr, const int32_t* indices, const float* data,
                       int64_t n_rows, const uint64_t* keys, int k, const char* names_blob, const int64_t* name_off,
                       const int64_t* row_len, int append);

/* deterministic synthetic polyploid genome (bench + tests).  See csrc/synth.cpp. */
typedef struct pc_synth_params {
    uint64_t seed;
    int64_t  genome_size;        /* total bases over all chromosomes */
    int32_t  n_subgenomes;       /* 2 = allotetraploid, 3 = hexaploid */
    int32_t  n_chrom;            /* chromosomes per subgenome */
    double   divergence;         /* homoeolog substitution divergence (0.08) */
    double   repeat_fraction;    /* target fraction of bases under repeats (0.60) */
    int32_t  n_shared_families;  /* 200 */
    int32_t  n_specific_families;/* 300 (split round-robin over subgenomes) */
    int32_t  min_family_len, max_family_len;   /* 500..8000 */
    double   max_copy_mutation;  /* U[0, 0.10] */
    double   n_fraction;         /* 0.001 */
    int32_t  min_n_run, max_n_run;             /* 100..10000 */
    double   softmask_fraction;  /* 0.10 */
} pc_synth_params;
int64_t pc_synth_n_chrom(const pc_synth_params* p);
int     pc_synth_chrom_len(const pc_synth_params* p, int64_t chrom, int64_t* len, char name[64]);
int     pc_synth_chrom(const pc_synth_params* p, int64_t chrom, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif
